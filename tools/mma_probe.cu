// mma_probe.cu -- microbenchmark behind the head kernel's design choices (not part of the product path).
// One CTA (or CTA pair) per SM issues back-to-back tcgen05.mma instructions on operands that sit in shared
// memory, optionally with tcgen05.commit every few MMAs and bulk-copy (TMA) traffic landing in shared
// memory at the same time, and reports cycles per MMA.
//   usage: mma_probe pair N a_mn per_commit tma n_mma [bk]
//     pair        0: cta_group::1 (M=128)   1: cta_group::2 (M=256 over two CTAs)
//     N           accumulator columns (multiple of 16)
//     a_mn        1: A is MN-major (pixels contiguous, as the NCHW feature maps), 0: K-major
//     per_commit  MMAs between two tcgen05.commit (0 = one commit at the end)
//     tma         0: none, 1: bulk loads from an L2-resident buffer, 2: from an HBM-sized buffer
//     n_mma       MMAs per CTA
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../yolo_continuous_b200/csrc/yc_sm100.cuh"

using namespace yc::sm100;

constexpr int A_BYTES = 128 * 64 * 2;   // 128 px x 64 k
constexpr int B_BYTES = 256 * 64 * 2;   // up to 256 rows x 64 k
constexpr int STAGE = A_BYTES + B_BYTES;
constexpr int NS = 3;
constexpr int SCRATCH = 4 * 16384;

struct Params {
    int pair, N, a_mn, per_commit, tma, n_mma, mode, fence, pollers, group;
    const uint8_t *src;
    size_t src_bytes;
    long long *cycles;   // per CTA
    long long *tma_bytes;
};

__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(512, 1) probe(const Params P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *scratch = smem + NS * STAGE;
    uint64_t *bars = (uint64_t *)(scratch + SCRATCH);
    uint64_t *cbar = bars;        // [8] commit barriers
    uint64_t *done = bars + 8;
    uint64_t *lbar = bars + 9;    // [4] load barriers
    volatile int *stop = (volatile int *)(bars + 13);
    uint32_t *tmem_ptr = (uint32_t *)(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (PAIR) rank = cluster_ctarank();

    // small finite bf16 values in the operand stages
    for (int i = threadIdx.x; i < NS * STAGE / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u + (i & 0xff);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&cbar[i], 1);
        mbar_init(done, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&lbar[i], 1);
        *stop = 0;
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair(tmem_ptr, 512);
        else tmem_alloc(tmem_ptr, 512);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t idesc = instr_desc_f16(1, (uint32_t)P.a_mn, 0, PAIR ? 256u : 128u, (uint32_t)P.N);

    if (P.mode == 1) {
        // commit latency: { per_commit MMAs; commit; wait } repeated n_mma times
        if (warp == 1 && lane == 0 && rank == 0) {
            const uint32_t sa0 = smem_addr(smem);
            const uint64_t da0 = smem_desc(sa0, A_BYTES / 2, 1024, SWZ_128B), db0 = smem_desc(sa0 + A_BYTES, 16, 1024, SWZ_128B);
            const long long t0 = clock64();
            uint32_t ph = 0;
            for (int i = 0; i < P.n_mma; ++i) {
                for (int k = 0; k < P.per_commit; ++k) {
                    if (PAIR) mma_f16_pair(tmem_base, da0, db0, idesc, 1u);
                    else mma_f16(tmem_base, da0, db0, idesc, 1u);
                }
                if (PAIR) mma_commit_pair(&cbar[0]);
                else mma_commit(&cbar[0]);
                mbar_wait(&cbar[0], ph);
                ph ^= 1u;
            }
            P.cycles[blockIdx.x] = clock64() - t0;
        }
    } else if (P.mode == 2 || P.mode == 3) {
        // ping-pong between two single threads in different warps: A arrives on cbar[0], B answers on cbar[1]
        // (mode 3: B answers with tcgen05.commit, as the MMA thread of the head kernel does)
        if (warp == 0 && lane == 0 && rank == 0) {
            const long long t0 = clock64();
            uint32_t ph = 0;
            for (int i = 0; i < P.n_mma; ++i) {
                mbar_arrive(&cbar[0]);
                mbar_wait(&cbar[1], ph);
                ph ^= 1u;
            }
            P.cycles[blockIdx.x] = clock64() - t0;
        } else if (warp == 1 && lane == 0 && rank == 0) {
            uint32_t ph = 0;
            for (int i = 0; i < P.n_mma; ++i) {
                mbar_wait(&cbar[0], ph);
                ph ^= 1u;
                if (P.mode == 3) {
                    tc_fence_after();
                    mma_commit(&cbar[1]);
                } else {
                    mbar_arrive(&cbar[1]);
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // whole warp runs the loop, one elected lane issues (the structure of the head kernel's MMA warp)
        const long long t0 = clock64();
        int c = 0;
        const uint32_t sa0 = smem_addr(smem);
        const uint64_t da0 = P.a_mn ? smem_desc(sa0, A_BYTES / 2, 1024, SWZ_128B) : smem_desc(sa0, 16, 1024, SWZ_128B);
        const uint64_t db0 = smem_desc(sa0 + A_BYTES, 16, 1024, SWZ_128B);
        const uint32_t kstep = P.a_mn ? 2048u >> 4 : 32u >> 4;
        const uint32_t cmask = P.per_commit ? (uint32_t)P.per_commit - 1u : 0xffffffffu; // per_commit: power of two
        const bool f_wait = P.fence & 2, f_fence = P.fence & 1;
        int st = 0;
        const bool f_one = P.fence & 4;
        const int G = P.group;
        for (int i = 0; i < P.n_mma; i += G) {
            const uint64_t so = (uint64_t)((uint32_t)(st * STAGE) >> 4);
            const uint32_t td = tmem_base + (uint32_t)((i >> 6) & 1) * 256u;
            if (f_wait) {   // a barrier whose phase-1 wait passes at once (fresh barrier)
                if (f_one) { if (lane == 0) mbar_wait(lbar + 3, 1u); __syncwarp(); }
                else mbar_wait(lbar + 3, 1u);
            }
            if (f_fence) tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (PAIR) mma_f16_pair(td, da0 + so + k * kstep, db0 + so + k * 2, idesc, (uint32_t)((i | k) != 0));
                    else mma_f16(td, da0 + so + k * kstep, db0 + so + k * 2, idesc, (uint32_t)((i | k) != 0));
                }
                if (G == 8) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (PAIR) mma_f16_pair(td, da0 + so + k * kstep, db0 + so + k * 2, idesc, 1u);
                        else mma_f16(td, da0 + so + k * kstep, db0 + so + k * 2, idesc, 1u);
                    }
                }
                if ((((uint32_t)i + (uint32_t)G) & cmask) == 0u) {
                    if (PAIR) mma_commit_pair(&cbar[c & 7]);
                    else mma_commit(&cbar[c & 7]);
                }
            }
            __syncwarp();
            ++c;
            if (++st == NS) st = 0;
        }
        if (elect_one()) {
            if (PAIR) mma_commit_pair(done);
            else mma_commit(done);
        }
        __syncwarp();
        mbar_wait(done, 0);
        const long long t1 = clock64();
        if (lane == 0) {
            P.cycles[blockIdx.x] = t1 - t0;
            *stop = 1;
            if (PAIR) asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(smem_addr((const void *)stop) | 0x01000000u), "r"(1) : "memory");
        }
    } else if (warp >= 4 && P.mode == 0) {
        // epilogue-like warps: every lane polls a barrier that never completes, until the MMA warp is done
        while (!*stop) {
            if (mbar_try_wait(lbar + 2, 0u)) break;
        }
    } else if (warp == 0 && lane == 0 && P.tma && P.mode == 0) {
        // bulk loads, 4 in flight, 16 KB each
        const size_t span = P.tma == 1 ? (size_t)(1 << 20) : (P.src_bytes / gridDim.x) & ~(size_t)1023;
        const uint8_t *base = P.src + (P.tma == 1 ? (size_t)blockIdx.x * span % (P.src_bytes - span) : (size_t)blockIdx.x * span);
        size_t off = 0;
        long long n = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        for (int s = 0; s < 4; ++s) {
            mbar_arrive_expect_tx(&lbar[s], 16384);
            bulk_load(scratch + s * 16384, base + off, 16384, &lbar[s]);
            off = (off + 16384) % (span - 16384);
            off &= ~(size_t)15;
        }
        while (!*stop) {
            for (int s = 0; s < 4; ++s) {
                mbar_wait(&lbar[s], ph[s]);
                ph[s] ^= 1;
                ++n;
                mbar_arrive_expect_tx(&lbar[s], 16384);
                bulk_load(scratch + s * 16384, base + off, 16384, &lbar[s]);
                off = (off + 16384) % (span - 16384);
                off &= ~(size_t)15;
            }
        }
        for (int s = 0; s < 4; ++s) mbar_wait(&lbar[s], ph[s]);
        P.tma_bytes[blockIdx.x] = n * 16384;
    }
    if (PAIR && rank == 1 && threadIdx.x == 0) P.cycles[blockIdx.x] = 0;
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    tc_fence_after();
    if (warp == 2) {
        if (PAIR) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

int main(int argc, char **argv)
{
    Params P;
    P.pair = argc > 1 ? atoi(argv[1]) : 0;
    P.N = argc > 2 ? atoi(argv[2]) : 256;
    P.a_mn = argc > 3 ? atoi(argv[3]) : 1;
    P.per_commit = argc > 4 ? atoi(argv[4]) : 4;
    P.tma = argc > 5 ? atoi(argv[5]) : 0;
    P.n_mma = argc > 6 ? atoi(argv[6]) : 4096;
    P.mode = argc > 7 ? atoi(argv[7]) : 0;
    P.fence = argc > 8 ? atoi(argv[8]) : 0;
    P.pollers = argc > 9 ? atoi(argv[9]) : 0;
    P.group = argc > 10 ? atoi(argv[10]) : 4;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = P.pair ? (sms / 2) * 2 : sms;
    P.src_bytes = (size_t)1 << 30;
    uint8_t *src;
    cudaMalloc(&src, P.src_bytes);
    cudaMemset(src, 0, P.src_bytes);
    P.src = src;
    cudaMalloc(&P.cycles, grid * sizeof(long long));
    cudaMalloc(&P.tma_bytes, grid * sizeof(long long));
    cudaMemset(P.tma_bytes, 0, grid * sizeof(long long));
    const size_t smem_bytes = 1024 + NS * STAGE + SCRATCH + 256;
    cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (P.pair) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(128 + 32 * P.pollers);
            cfg.dynamicSmemBytes = smem_bytes;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, probe<true>, P);
        } else {
            probe<false><<<grid, 128 + 32 * P.pollers, smem_bytes>>>(P);
        }
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long *h = (long long *)malloc(grid * sizeof(long long)), *hb = (long long *)malloc(grid * sizeof(long long));
    cudaMemcpy(h, P.cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hb, P.tma_bytes, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mn = 1LL << 60, mx = 0, sum = 0, tb = 0;
    int cnt = 0;
    for (int i = 0; i < grid; ++i) {
        tb += hb[i];
        if (!h[i]) continue;
        ++cnt; sum += h[i];
        if (h[i] < mn) mn = h[i];
        if (h[i] > mx) mx = h[i];
    }
    const double flops = 2.0 * (P.pair ? 256 : 128) * P.N * 16 * (double)P.n_mma * cnt;
    if (P.mode) {
        printf("mode=%d pair=%d N=%d mmas_per_commit=%d iters=%d : cycles per iteration avg %.1f min %.1f max %.1f\n", P.mode, P.pair,
               P.N, P.per_commit, P.n_mma, (double)sum / cnt / P.n_mma, (double)mn / P.n_mma, (double)mx / P.n_mma);
        return 0;
    }
    printf("fence=%d pollers=%d group=%d ", P.fence, P.pollers, P.group);
    printf("pair=%d N=%d a_mn=%d per_commit=%d tma=%d n_mma=%d : cyc/MMA avg %.1f min %.1f max %.1f | %.3f ms, %.0f TFLOP/s, "
           "tma %.1f GB/s (%.1f B/clk/SM)\n",
           P.pair, P.N, P.a_mn, P.per_commit, P.tma, P.n_mma, (double)sum / cnt / P.n_mma, (double)mn / P.n_mma,
           (double)mx / P.n_mma, ms, flops / (ms * 1e-3) / 1e12, tb / (ms * 1e-3) / 1e9,
           (double)tb / grid / ((double)sum / cnt));
    return 0;
}
