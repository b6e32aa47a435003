// yc_head_generic.cu -- parameter packing and the any-shape head path.
//
//  * yc_head_pack: folds ImplicitA into the bias (reference nets/common.py:425-426 applied before the
//    conv at nets/idetect.py:31), keeps ImplicitM (nets/common.py:438-439) as an epilogue scale and
//    produces the operand formats of the tcgen05 kernels (bf16 for bf16 maps; a per-row scaled fp16 hi/lo split,
//    transposed, for float32 maps: yc_head_sm100_split.cu).
//  * head_generic_kernel: FFMA-tiled 1x1 conv with exact binary32 accumulation and the decode fused in
//    the epilogue.  It is the path for shapes the TMA/tcgen05 kernel cannot take (feature-map rows
//    not 16-byte multiples, na*no > 256, K % 16 != 0) and the on-device cross-check of that kernel.
//  * ibin_decode_kernel, decode_box_kernel: row decoders that need whole rows (IBin argmax over bins,
//    Variant A NCHW input).
#include "yc_common.cuh"

namespace yc {

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) head_pack_kernel(const float *__restrict__ W, const float *__restrict__ bias,
                                                        const float *__restrict__ ia, const float *__restrict__ im, int N,
                                                        int K, int Npad, int no, int npad_g, int wt, float *bias2, float *scale, float *scale_split,
                                                        float2 *sb, float2 *sb_split, float *w32, __half *w_hi, __half *w_lo, __nv_bfloat16 *w_bf)
{
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= Npad) return;
    if (c >= N) { // zero padding rows: the MMA sees zeros, the epilogue never reads them
        for (int k = lane; k < K; k += 32) w_bf[(size_t)c * K + k] = __float2bfloat16_rn(0.f);
        if (lane == 0) {
            bias2[c] = 0.f; scale[c] = 0.f; scale_split[c] = 0.f;
            sb[c] = make_float2(0.f, 0.f); sb_split[c] = make_float2(0.f, 0.f);
        }
        return;
    }
    const float *wr = W + (size_t)c * K;
    double dot = 0.0;
    float amax = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = wr[k];
        amax = fmaxf(amax, fabsf(w));
        if (ia) dot += (double)w * (double)ia[k];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, d);
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
    }
    // scale the row so that its largest weight lands in [2^12, 2^13): the fp16 "lo" parts of all
    // weights within 2^-15 of the largest stay normal numbers
    int shift = 0;
    if (amax > 0.f && isfinite(amax)) shift = 12 - ilogbf(amax);
    shift = max(-60, min(60, shift));
    const float up = ldexpf(1.0f, shift);
    const int tcol = (c / no) * npad_g + c % no;   // column of channel c in the per-anchor padded transposed copies
    for (int k = lane; k < K; k += 32) {
        const float w = wr[k];
        const float ws = w * up; // exact (power of two)
        const __half hi = __float2half_rn(ws);
        const __half lo = __float2half_rn(ws - __half2float(hi));
        w_hi[(size_t)k * wt + tcol] = hi;   // transposed: the split kernel reads the weights as an MN-major operand
        w_lo[(size_t)k * wt + tcol] = lo;
        w_bf[(size_t)c * K + k] = __float2bfloat16_rn(w);
        w32[(size_t)c * K + k] = w;
    }
    if (lane == 0) {
        const float b1 = (float)((double)(bias ? bias[c] : 0.f) + dot);
        const float m = im ? im[c] : 1.0f;
        bias2[c] = __fmul_rn(m, b1);
        scale[c] = m;
        scale_split[c] = ldexpf(m, YC_SPLIT_XSHIFT - shift);
        sb[c] = make_float2(m, __fmul_rn(m, b1));
        sb_split[c] = make_float2(ldexpf(m, YC_SPLIT_XSHIFT - shift), __fmul_rn(m, b1));
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

struct GenericLevel {
    const void *x;
    const void *w;      // w32 (float) or w_bf (bf16), [N or Npad, K]
    const float *bias2;
    const float *scale;
    float *raw;         // [bs, na, HW, no] or null
    float *z;           // base of z, or null
    int K, HW, nx;
    int row_off;        // first z row of this level
    float stride, stride_y;
    float anchor_wh[YC_MAX_ANCHORS * 2];
};

constexpr int GM = 128, GN = 64, GK = 16;

// grid: (ceil(HW/128), ceil(N/64), bs)   block: 256 threads, 8 pixels x 4 channels each.
// Exact binary32: every output is one fmaf chain over k in ascending order (what the oracle computes), whatever the
// tiling.  Operands come out of shared memory as float4 (32 FMAs per 3 loads) and the next k-tile is fetched from global
// memory into registers while the current one is multiplied.
template <typename XT, typename WT>
__global__ void __launch_bounds__(256) head_generic_kernel(GenericLevel L, int na, int no, int rows_total, int decode)
{
    __shared__ __align__(16) float As[GK][GM];
    __shared__ __align__(16) float Bs[GK][GN + 4];   // +4: the transposing stores below hit 8 banks instead of 1
    const int N = na * no;
    const int p0 = blockIdx.x * GM, c0 = blockIdx.y * GN, b = blockIdx.z;
    // tx -> pixels tx*4..+3 and 64+tx*4..+3 (two conflict-free float4 reads), ty -> channels ty*4..+3
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const XT *x = (const XT *)L.x + (size_t)b * L.K * L.HW;
    const WT *w = (const WT *)L.w;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[8], rb[4];   // the next k-tile's global loads: A 16 x 128 (coalesced along pixels), B 64 x 16 (along k)
#define YC_GENERIC_FETCH(K0)                                                                          \
    {                                                                                                 \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) {                                               \
            const int e = tid + i * 256, kk = e >> 7, pp = e & 127;                                   \
            const int k = (K0) + kk, p = p0 + pp;                                                     \
            ra[i] = (k < L.K && p < L.HW) ? to_f32<XT>(x[(size_t)k * L.HW + p]) : 0.f;                \
        }                                                                                             \
        _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                               \
            const int e = tid + i * 256, cc = e >> 4, kk = e & 15;                                    \
            const int k = (K0) + kk, c = c0 + cc;                                                     \
            rb[i] = (k < L.K && c < N) ? to_f32<WT>(w[(size_t)c * L.K + k]) : 0.f;                    \
        }                                                                                             \
    }
    YC_GENERIC_FETCH(0)
    for (int k0 = 0; k0 < L.K; k0 += GK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            As[e >> 7][e & 127] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            Bs[e & 15][e >> 4] = rb[i];
        }
        __syncthreads();
        if (k0 + GK < L.K) YC_GENERIC_FETCH(k0 + GK)
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *(const float4 *)&As[kk][tx * 4], a1 = *(const float4 *)&As[kk][64 + tx * 4];
            const float4 b4 = *(const float4 *)&Bs[kk][ty * 4];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty * 4 + j;
        if (c >= N) continue;
        const int a = c / no, o = c - a * no;
        const float sc = L.scale[c], bi = L.bias2[c];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = p0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + i - 4);
            if (p >= L.HW) continue;
            const float t = fmaf(acc[i][j], sc, bi);
            if (L.raw) L.raw[(((size_t)b * na + a) * L.HW + p) * no + o] = t;
            if (decode) {
                float s = sigmoidf_fast(t);
                if (o == 0) s = decode_xy(s, (float)(p % L.nx), L.stride);
                else if (o == 1) s = decode_xy(s, (float)(p / L.nx), L.stride_y);
                else if (o < 4) s = decode_wh(s, L.anchor_wh[a * 2 + (o - 2)]);
                L.z[((size_t)b * rows_total + L.row_off + (size_t)a * L.HW + p) * no + o] = s;
            }
        }
    }
#undef YC_GENERIC_FETCH
}

// ------------------------------------------------------------------------------------------------
// IBin decode (reference nets/ibin.py:56-72, losses/sigmoid_bin.py:49-63): one thread per output element.
__global__ void __launch_bounds__(256) ibin_decode_kernel(const float *__restrict__ raw, int bs, int na, int HW, int nx,
                                                          int no_in, int bin_count, float stride, float stride_y, float step,
                                                          const float *__restrict__ bins, float a_w0, float a_h0,
                                                          float a_w1, float a_h1, float a_w2, float a_h2, float a_w3,
                                                          float a_h3, float *__restrict__ z, int rows_total, int row_off)
{
    const int len = bin_count + 1, no_out = no_in - 2 * len + 2;
    const size_t total = (size_t)bs * na * HW * no_out;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int j = (int)(i % no_out);
    const size_t r = i / no_out;
    const int p = (int)(r % HW);
    const int a = (int)((r / HW) % na);
    const int b = (int)(r / ((size_t)HW * na));
    const float *q = raw + r * no_in;
    float v;
    if (j < 2) {
        v = j == 0 ? decode_xy(sigmoidf_fast(q[j]), (float)(p % nx), stride) : decode_xy(sigmoidf_fast(q[j]), (float)(p / nx), stride_y);
    } else if (j < 4) {
        const float *g = q + 2 + (j - 2) * len;
        float reg = __fmul_rn(sigmoidf_fast(g[0]), 2.0f);
        reg = __fadd_rn(reg, -1.0f);
        reg = __fmul_rn(reg, step);
        int best = 0;
        float bv = sigmoidf_fast(g[1]);
        for (int k = 1; k < bin_count; ++k) {
            const float sv = sigmoidf_fast(g[1 + k]);
            if (sv > bv) { bv = sv; best = k; }
        }
        float res = __fadd_rn(reg, bins[best]);
        res = fminf(fmaxf(res, 0.0f), 4.0f);
        const float aw[8] = {a_w0, a_h0, a_w1, a_h1, a_w2, a_h2, a_w3, a_h3};
        v = __fmul_rn(res, aw[a * 2 + (j - 2)]);
    } else {
        v = sigmoidf_fast(q[2 + 2 * len + (j - 4)]);
    }
    z[((size_t)b * rows_total + row_off + (size_t)a * HW + p) * no_out + j] = v;
}

// Variant A decode (reference detect.py:29-87): conv NCHW -> normalised rows.
__device__ __forceinline__ float decode_box_value(float t, int j, int p, int a, int ny, int nx, const float *aw)
{
    float s = sigmoidf_fast(t);
    if (j < 2) {
        float v = __fmul_rn(s, 2.0f);
        v = __fadd_rn(v, -0.5f);
        v = __fadd_rn(v, j == 0 ? (float)(p % nx) : (float)(p / nx));
        s = __fdiv_rn(v, j == 0 ? (float)nx : (float)ny);
    } else if (j < 4) {
        float v = __fmul_rn(s, 2.0f);
        v = __fmul_rn(v, v);
        v = __fmul_rn(v, aw[a * 2 + (j - 2)]);
        s = __fdiv_rn(v, j == 2 ? (float)nx : (float)ny);
    }
    return s;
}

struct Anchors8 { float v[8]; };

// The map is channel-major ([no][HW] per anchor), the rows are channel-minor ([HW][no]): a CTA transposes a
// [no][TP pixels] tile through shared memory, so that both the reads (along pixels) and the writes (TP consecutive rows
// = one contiguous span of the output) are coalesced.  grid (ceil(HW/TP), bs*na).
template <int TP>
__global__ void __launch_bounds__(256) decode_box_kernel(const float *__restrict__ conv, int na, int no, int ny, int nx,
                                                         Anchors8 aw, float *__restrict__ out)
{
    extern __shared__ float tile[];   // [no][TP + 1]
    const int HW = ny * nx, p0 = blockIdx.x * TP;
    const size_t ba = blockIdx.y;     // b*na + a
    const int a = (int)(ba % na);
    const int np = min(TP, HW - p0);
    const float *src = conv + ba * no * (size_t)HW;
    for (int i = threadIdx.x; i < no * TP; i += 256) {
        const int j = i / TP, pp = i - j * TP;
        if (pp < np) tile[j * (TP + 1) + pp] = decode_box_value(src[(size_t)j * HW + p0 + pp], j, p0 + pp, a, ny, nx, aw.v);
    }
    __syncthreads();
    float *dst = out + (ba * HW + p0) * no;
    for (int i = threadIdx.x; i < np * no; i += 256) {
        const int pp = i / no, j = i - pp * no;
        dst[i] = tile[j * (TP + 1) + pp];
    }
}

// host launcher used by yc_head_forward (yc_abi.cu)
int launch_head_generic(const yc_head_desc *d, int rows_total, const int *row_off, unsigned level_mask,
                        cudaStream_t stream)
{
    const int N = d->na * d->no;
    for (int i = 0; i < d->nl; ++i) {
        if (!(level_mask >> i & 1u)) continue;
        const yc_head_level &lv = d->level[i];
        const int HW = lv.H * lv.W;
        BlobView bv = blob_view(lv.blob, N, lv.K);
        GenericLevel L;
        L.x = lv.x;
        L.w = d->x_dtype == YC_BF16 ? (const void *)bv.w_bf : (const void *)bv.w32;
        L.bias2 = bv.bias2;
        L.scale = bv.scale;
        L.raw = lv.raw;
        L.z = d->z;
        L.K = lv.K; L.HW = HW; L.nx = lv.W;
        L.row_off = row_off[i];
        L.stride = lv.stride;
        L.stride_y = lv.stride_y > 0.f ? lv.stride_y : lv.stride;
        for (int j = 0; j < YC_MAX_ANCHORS * 2; ++j) L.anchor_wh[j] = lv.anchor_wh[j];
        const int decode = d->kind == YC_HEAD_IDETECT ? 1 : 0;
        if (d->kind != YC_HEAD_IDETECT)
            YC_REQUIRE(lv.raw != nullptr, YC_ERR_INVALID, "generic head path: kind %d needs raw buffers", d->kind);
        dim3 grid((HW + GM - 1) / GM, (N + GN - 1) / GN, d->bs);
        if (d->x_dtype == YC_BF16)
            head_generic_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(L, d->na, d->no, rows_total, decode);
        else
            head_generic_kernel<float, float><<<grid, 256, 0, stream>>>(L, d->na, d->no, rows_total, decode);
        if (d->kind == YC_HEAD_IBIN) {
            const int len = d->bin_count + 1, no_out = d->no - 2 * len + 2;
            const size_t total = (size_t)d->bs * d->na * HW * no_out;
            const float *a = lv.anchor_wh;
            ibin_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
                lv.raw, d->bs, d->na, HW, lv.W, d->no, d->bin_count, lv.stride, lv.stride_y > 0.f ? lv.stride_y : lv.stride, (float)(4.0 / (double)d->bin_count), d->bins, a[0], a[1], a[2], a[3], a[4],
                a[5], a[6], a[7], d->z, rows_total, row_off[i]);
        }
    }
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc

using namespace yc;

extern "C" size_t yc_head_pack_bytes(int N, int K)
{
    if (N <= 0 || K <= 0) return 0;
    const int Npad = round_up(N, 16);
    size_t w32b = round_up_sz(sizeof(float) * (size_t)N * K, 128);
    size_t w16b = round_up_sz(sizeof(__half) * (size_t)Npad * K, 128);
    size_t wtb = round_up_sz(sizeof(__half) * (size_t)(Npad + 16 * YC_MAX_ANCHORS) * K, 128);
    return sizeof(float) * 7 * (size_t)Npad + w32b + w16b + 2 * wtb + 256;
}

extern "C" int yc_head_pack(const float *W, const float *bias, const float *ia, const float *im, int N, int K, int na,
                            void *blob, yc_stream_t stream)
{
    YC_REQUIRE(W && blob && N > 0 && K > 0, YC_ERR_INVALID, "yc_head_pack: bad argument");
    YC_REQUIRE(na >= 1 && na <= YC_MAX_ANCHORS && N % na == 0, YC_ERR_INVALID, "yc_head_pack: N = %d is not %d anchors x no", N, na);
    YC_REQUIRE(((uintptr_t)blob & 127) == 0, YC_ERR_INVALID, "yc_head_pack: blob must be 128-byte aligned");
    const int Npad = round_up(N, 16);
    const int no = N / na, npad_g = round_up(no, 16), wt = na * npad_g;
    BlobView v = blob_view(blob, N, K);
    // padding columns of the transposed copies feed accumulator columns nobody reads, but must not hold NaN patterns
    YC_CUDA(cudaMemsetAsync((void *)v.w_hi_t, 0, (size_t)((const char *)v.w_bf - (const char *)v.w_hi_t), (cudaStream_t)stream));
    head_pack_kernel<<<(Npad + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
        W, bias, ia, im, N, K, Npad, no, npad_g, wt, (float *)v.bias2, (float *)v.scale, (float *)v.scale_split, (float2 *)v.sb, (float2 *)v.sb_split,
        (float *)v.w32,
        (__half *)v.w_hi_t, (__half *)v.w_lo_t, (__nv_bfloat16 *)v.w_bf);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

extern "C" int yc_decode_box(const float *conv, int bs, int na, int no, int ny, int nx,
                             const float *anchor_wh_scaled_host, float *out, yc_stream_t stream)
{
    YC_REQUIRE(conv && out && anchor_wh_scaled_host, YC_ERR_INVALID, "yc_decode_box: null argument");
    YC_REQUIRE(bs > 0 && na > 0 && na <= YC_MAX_ANCHORS && no >= 5 && ny > 0 && nx > 0, YC_ERR_INVALID,
               "yc_decode_box: bad shape");
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < na * 2; ++i) a[i] = anchor_wh_scaled_host[i];
    Anchors8 aw;
    for (int i = 0; i < 8; ++i) aw.v[i] = a[i];
    const int HW = ny * nx;
    YC_REQUIRE((size_t)bs * na <= 65535, YC_ERR_UNSUPPORTED, "yc_decode_box: bs*na > 65535");
    if ((size_t)no * 65 * 4 <= 48 * 1024) {
        decode_box_kernel<64><<<dim3((HW + 63) / 64, bs * na), 256, (size_t)no * 65 * 4, (cudaStream_t)stream>>>(
            conv, na, no, ny, nx, aw, out);
    } else {
        YC_REQUIRE((size_t)no * 9 * 4 <= 48 * 1024, YC_ERR_UNSUPPORTED, "yc_decode_box: no = %d too large", no);
        decode_box_kernel<8><<<dim3((HW + 7) / 8, bs * na), 256, (size_t)no * 9 * 4, (cudaStream_t)stream>>>(
            conv, na, no, ny, nx, aw, out);
    }
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}
