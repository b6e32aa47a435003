// yc_exchange.cu -- the one exchange of the multi-GPU path: every rank's detections of a step
// reach every other rank.  No collective library, no host work in the steady state: a rank's `push` kernel stores its
// fixed-size message (header + the first rows of its detection list) straight into a slot of every peer's receive buffer
// over NVLink (peer-mapped memory obtained through CUDA IPC), fences at system scope and raises a per-(slot, source) flag
// in the peer's buffer; the `wait` kernel of the consumer polls its own flags.  Sequence numbers live in device memory, so
// both kernels can sit inside a captured CUDA graph that is replayed every step.
//
// Flow control (credits): a message with sequence number q goes to slot q % slots.  Before it overwrites that slot in
// peer p, the sender waits until p has acknowledged q - slots, i.e. until p's wait kernel for sequence number
// q - slots + 1 has run (calling wait(j) declares everything before j consumed).  Acknowledgements are remote stores
// into the sender's buffer, like the flags.  Every spin is bounded by a timeout that sets a sticky error word.
//
// Receive buffer of a rank (one cudaMalloc, exported with cudaIpcGetMemHandle):
//   [slots][world][msg_bytes] messages | flags [slots][world] u32 (seq + 1 of the message in the slot)
//   | acks [world] u32 (acks[p]: peer p has consumed every message before this sequence number) | state (local only)
#include "yc_common.cuh"

namespace yc {

struct XchgState {             // local to a rank (lives behind its receive buffer)
    unsigned int seq_push;     // sequence number of the next message this rank pushes
    unsigned int seq_wait;     // sequence number of the next message set this rank waits for
    unsigned int done;         // CTAs of the running push kernel that have finished
    unsigned int error;        // sticky: 1 = a wait timed out
    unsigned int peer_done[60]; // CTAs of the running push kernel that have finished their part for peer p
};
constexpr int XCHG_MAX_WORLD = 60;
constexpr int XCHG_SPLIT = 4;  // CTAs per peer in the push kernel
constexpr int XCHG_UNROLL = 8; // 16-byte loads in flight per thread: 4 CTAs x 256 threads x 8 x 16 B = 128 KB per peer

struct XchgLayout {
    size_t flags_off, acks_off, state_off, total;
};

static inline XchgLayout xchg_layout(int world, int slots, size_t msg_bytes)
{
    XchgLayout l;
    size_t p = round_up_sz((size_t)slots * world * msg_bytes, 256);
    l.flags_off = p; p += round_up_sz((size_t)slots * world * 4, 256);
    l.acks_off = p;  p += round_up_sz((size_t)world * 4, 256);
    l.state_off = p; p += 256;
    l.total = p;
    return l;
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long XCHG_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;

// grid = world x XCHG_SPLIT CTAs: the CTAs of peer p copy this rank's message into peer p's slot; the last of them raises
// the flag there.
// 256 threads x <= 48 registers: like the NMS kernels, a CTA must fit NEXT TO a resident head CTA (512 threads x 96
// registers, 213 KB of shared memory) -- the push of step i runs while the persistent head kernel of step i + 1 holds
// every SM; a CTA that does not fit would wait for that kernel to end and stall the pipeline by a step.
// msg: [hdr_ints int32: counts (bs) | offsets (bs + 1) | pad][rows: 7 floats each]; only the rows that exist (and fit) move.
__global__ void __maxnreg__(48) xchg_push_kernel(const uint8_t *__restrict__ msg, int hdr_ints, int bs, int max_rows,
                                                        uint8_t *const *__restrict__ peers, int world, int rank, int slots,
                                                        size_t msg_bytes, XchgLayout lay)
{
    const int p = blockIdx.x / XCHG_SPLIT, part = blockIdx.x % XCHG_SPLIT;
    uint8_t *mine = peers[rank];
    XchgState *st = (XchgState *)(mine + lay.state_off);
    __shared__ unsigned int s_q;
    if (threadIdx.x == 0) {
        const unsigned int q = *(volatile unsigned int *)&st->seq_push;
        // credit: peer p has consumed the message that last used this slot
        if (q >= (unsigned)slots) {
            const unsigned int *ack = (const unsigned int *)(mine + lay.acks_off) + p;
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(ack) + (unsigned)slots <= q) {
                if (global_ns() - t0 > XCHG_TIMEOUT_NS) { atomicExch(&st->error, 1u); break; }
                __nanosleep(200);
            }
        }
        s_q = q;
    }
    __syncthreads();
    const unsigned int q = s_q;
    const int slot = (int)(q % (unsigned)slots);
    const int total = ((const int *)msg)[2 * bs];
    const size_t n16 = ((size_t)hdr_ints * 4 + (size_t)min(total, max_rows) * 28 + 15) / 16;
    uint8_t *dst = peers[p] + ((size_t)slot * world + rank) * msg_bytes;
    const uint4 *s4 = (const uint4 *)msg;
    uint4 *d4 = (uint4 *)dst;
    // next to a head kernel that saturates HBM a dependent load -> store round trip takes microseconds: all loads of a
    // thread are issued before its first store (the grid is sized so that XCHG_UNROLL iterations cover a full message)
    const size_t lo = n16 * part / XCHG_SPLIT, hi = n16 * (part + 1) / XCHG_SPLIT;
    for (size_t i0 = lo + threadIdx.x; i0 < hi; i0 += (size_t)blockDim.x * XCHG_UNROLL) {
        uint4 v[XCHG_UNROLL];
#pragma unroll
        for (int u = 0; u < XCHG_UNROLL; ++u)
            if (i0 + (size_t)u * blockDim.x < hi) v[u] = __ldcs(s4 + i0 + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < XCHG_UNROLL; ++u)
            if (i0 + (size_t)u * blockDim.x < hi) d4[i0 + (size_t)u * blockDim.x] = v[u];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // ONE system-scope fence per CTA (after the barrier it orders the stores of all its threads): a fence per thread
        // next to a head kernel with thousands of loads in flight cost ~20 us
        __threadfence_system();
        if (atomicAdd(&st->peer_done[p], 1u) == XCHG_SPLIT - 1) {   // all parts for this peer are on their way: raise the flag
            st->peer_done[p] = 0u;
            st_release_sys((unsigned int *)(peers[p] + lay.flags_off) + (size_t)slot * world + rank, q + 1u);
        }
        __threadfence();
        if (atomicAdd(&st->done, 1u) == (unsigned)(world * XCHG_SPLIT) - 1u) {   // last CTA of the kernel
            st->done = 0u;
            __threadfence();
            *(volatile unsigned int *)&st->seq_push = q + 1u;
        }
    }
}

// one CTA: thread r waits for rank r's message with this rank's next sequence number j, then acknowledges to rank r that
// everything before j has been consumed here.  `lag`: do nothing unless this rank has itself pushed message j + lag
// already (lag = 1 in the steady state: the wait for step i - 1 sits behind the push of step i and practically never
// spins, while ranks still cannot drift more than a step apart).
__global__ void __launch_bounds__(32) xchg_wait_kernel(uint8_t *const *__restrict__ peers, int world, int rank, int slots,
                                                       int lag, XchgLayout lay)
{
    uint8_t *mine = peers[rank];
    XchgState *st = (XchgState *)(mine + lay.state_off);
    do {   // lag < 0: until every message this rank has pushed so far has been awaited
        const unsigned int j = *(volatile unsigned int *)&st->seq_wait;
        if (*(volatile unsigned int *)&st->seq_push < j + 1u + (unsigned)max(lag, 0)) return;
        const int slot = (int)(j % (unsigned)slots);
        for (int r = threadIdx.x; r < world; r += 32) {
            const unsigned int *flag = (const unsigned int *)(mine + lay.flags_off) + (size_t)slot * world + r;
            const unsigned long long t0 = global_ns();
            while ((int)(ld_acquire_sys(flag) - (j + 1u)) < 0) {
                if (global_ns() - t0 > XCHG_TIMEOUT_NS) { atomicExch(&st->error, 1u); break; }
                __nanosleep(200);
            }
            // the acknowledgement orders nothing of this kernel's: what it declares consumed was read by earlier work of
            // the stream, complete before this kernel started -- a plain (relaxed, system-scope) store
            *(volatile unsigned int *)((unsigned int *)(peers[r] + lay.acks_off) + rank) = j;
        }
        __syncwarp();
        if (threadIdx.x == 0) *(volatile unsigned int *)&st->seq_wait = j + 1u;
        __syncwarp();
    } while (lag < 0);
}

} // namespace yc

using namespace yc;

extern "C" size_t yc_xchg_bytes(int world, int slots, size_t msg_bytes)
{
    if (world <= 0 || slots <= 0 || msg_bytes == 0) return 0;
    return xchg_layout(world, slots, msg_bytes).total;
}

extern "C" int yc_xchg_alloc(int world, int slots, size_t msg_bytes, void **buf, uint8_t *handle64)
{
    static_assert(sizeof(XchgState) <= 256, "state block");
    YC_REQUIRE(world <= XCHG_MAX_WORLD, YC_ERR_UNSUPPORTED, "yc_xchg_alloc: at most %d ranks", XCHG_MAX_WORLD);
    YC_REQUIRE(buf && handle64 && world > 0 && slots >= 2 && msg_bytes % 16 == 0 && msg_bytes > 0, YC_ERR_INVALID,
               "yc_xchg_alloc: bad argument (msg_bytes must be a multiple of 16, slots >= 2)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const XchgLayout lay = xchg_layout(world, slots, msg_bytes);
    void *p = nullptr;
    YC_CUDA(cudaMalloc(&p, lay.total));
    YC_CUDA(cudaMemset(p, 0, lay.total));
    YC_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    YC_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *buf = p;
    return YC_OK;
}

extern "C" int yc_xchg_open(const uint8_t *handle64, void **peer_buf)
{
    YC_REQUIRE(handle64 && peer_buf, YC_ERR_INVALID, "yc_xchg_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    YC_CUDA(cudaIpcOpenMemHandle(peer_buf, h, cudaIpcMemLazyEnablePeerAccess));
    return YC_OK;
}

extern "C" int yc_xchg_close(void *peer_buf)
{
    if (peer_buf) YC_CUDA(cudaIpcCloseMemHandle(peer_buf));
    return YC_OK;
}

extern "C" int yc_xchg_free(void *buf)
{
    if (buf) YC_CUDA(cudaFree(buf));
    return YC_OK;
}

extern "C" int yc_xchg_push(const void *msg, int hdr_ints, int bs, int max_rows, void *const *peers_dev, int world, int rank,
                            int slots, size_t msg_bytes, yc_stream_t stream)
{
    YC_REQUIRE(msg && peers_dev && world > 0 && rank >= 0 && rank < world && slots >= 2 && bs > 0 && hdr_ints >= 2 * bs + 1,
               YC_ERR_INVALID, "yc_xchg_push: bad argument");
    YC_REQUIRE((size_t)hdr_ints * 4 + (size_t)max_rows * 28 <= msg_bytes && (((uintptr_t)msg | msg_bytes) & 15) == 0, YC_ERR_INVALID,
               "yc_xchg_push: message does not fit msg_bytes / is not 16-byte aligned");
    xchg_push_kernel<<<world * XCHG_SPLIT, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)msg, hdr_ints, bs, max_rows,
                                                               (uint8_t *const *)peers_dev, world, rank, slots, msg_bytes,
                                                               xchg_layout(world, slots, msg_bytes));
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

extern "C" int yc_xchg_wait(void *const *peers_dev, int world, int rank, int slots, size_t msg_bytes, int lag,
                            yc_stream_t stream)
{
    YC_REQUIRE(peers_dev && world > 0 && rank >= 0 && rank < world && slots >= 2 && lag >= -1 && lag < slots - 1, YC_ERR_INVALID,
               "yc_xchg_wait: bad argument");
    xchg_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((uint8_t *const *)peers_dev, world, rank, slots, lag,
                                                         xchg_layout(world, slots, msg_bytes));
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

// host-side view of the local state words (seq_push, seq_wait, done, error); synchronises `stream`
extern "C" int yc_xchg_state(const void *buf, int world, int slots, size_t msg_bytes, uint32_t *out4, yc_stream_t stream)
{
    YC_REQUIRE(buf && out4, YC_ERR_INVALID, "yc_xchg_state: null argument");
    const XchgLayout lay = xchg_layout(world, slots, msg_bytes);
    YC_CUDA(cudaMemcpyAsync(out4, (const uint8_t *)buf + lay.state_off, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    YC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return YC_OK;
}
