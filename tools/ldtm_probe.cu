// tools/ldtm_probe.cu -- what does tcgen05.ld.16x32bx2 return, and may its base lane be 16 within the warp's quadrant?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp/ldtm_probe tools/ldtm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t *out)
{
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_ptr;
    // fill: lane L (= 32*warp + lane), column c holds L * 1000 + c
    const uint32_t taddr = base + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < 128; ++c) {
        uint32_t v = (uint32_t)((32 * warp + lane) * 1000 + c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    // read back: 16x32bx2, x4, half-split offset 48, base lanes 0 and 16 of the warp's quadrant
    for (int half = 0; half < 2; ++half) {
        uint32_t r[4];
        const uint32_t ta = base + ((uint32_t)(32 * warp + 16 * half) << 16) + 5u;
        asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0, %1, %2, %3}, [%4], 48;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 4; ++j) out[((warp * 2 + half) * 32 + lane) * 4 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base));
}

int main()
{
    uint32_t *d, h[4 * 2 * 32 * 4];
    cudaMalloc(&d, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int w = 0; w < 4; w += 3)
        for (int half = 0; half < 2; ++half) {
            printf("warp %d base lane +%d (expect thread t<16: lane 32w+16h+t cols 5..8; t>=16: lane 32w+16h+t-16 cols 53..56)\n", w, 16 * half);
            for (int t = 0; t < 32; t += 5)
                printf("  t=%2d: %u %u %u %u\n", t, h[((w * 2 + half) * 32 + t) * 4], h[((w * 2 + half) * 32 + t) * 4 + 1],
                       h[((w * 2 + half) * 32 + t) * 4 + 2], h[((w * 2 + half) * 32 + t) * 4 + 3]);
        }
    return 0;
}
