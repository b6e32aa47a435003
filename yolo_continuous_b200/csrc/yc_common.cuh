// yc_common.cuh -- shared helpers for the sm_100a kernels of the detection post-backbone path.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/yc_b200.h"

namespace yc {

// ---- error plumbing (thread-local text behind yc_last_error) -----------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define YC_CUDA(call)                                              \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return yc::cuda_fail(_e, #call);    \
    } while (0)

#define YC_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            yc::set_error(__VA_ARGS__);      \
            return (code);                   \
        }                                    \
    } while (0)

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline size_t round_up_sz(size_t a, size_t b) { return (a + b - 1) / b * b; }

// float32 maps on the tensor cores (yc_head_sm100_split.cu): activations are scaled by 2^-YC_SPLIT_XSHIFT before they are
// split into two fp16 numbers (range +-65504 * 2^XSHIFT ~ 1e6; the low part of |x| < 2 is a subnormal, absolute error
// <= 2^-25 * 2^XSHIFT); the epilogue scale carries the inverse
#define YC_SPLIT_XSHIFT 4

// ---- blob layout produced by yc_head_pack (see include/yc_b200.h) ------------------------
struct BlobView {
    const float *bias2;       // im * (b + W.ia)
    const float *scale;       // im
    const float *scale_split; // im * 2^(YC_SPLIT_XSHIFT - shift(c))   (fp16 hi/lo path)
    const float2 *sb;         // (scale, bias2) interleaved: one 8-byte broadcast load per column in the epilogue
    const float2 *sb_split;   // (scale_split, bias2)
    const float *w32;         // [N,K]
    const __half *w_hi_t;     // [K, na*npad_g]  fp16(W * 2^shift(c)), transposed (MN-major B operand); anchor a's `no` channels
                              //   start at column a*npad_g, npad_g = round_up(no, 16): TMA box starts must be 16-byte aligned
    const __half *w_lo_t;     // [K, na*npad_g]  fp16(W * 2^shift(c) - w_hi)
    const __nv_bfloat16 *w_bf; // [Npad,K]
};

static inline size_t blob_off_scale(int Npad) { return sizeof(float) * (size_t)Npad; }

__host__ __device__ inline BlobView blob_view(const void *blob, int N, int K)
{
    const int Npad = (N + 15) / 16 * 16;
    const char *p = (const char *)blob;
    BlobView v;
    v.bias2 = (const float *)p;                  p += sizeof(float) * (size_t)Npad;
    v.scale = (const float *)p;                  p += sizeof(float) * (size_t)Npad;
    v.scale_split = (const float *)p;            p += sizeof(float) * (size_t)Npad;
    v.sb = (const float2 *)p;                    p += sizeof(float2) * (size_t)Npad;
    v.sb_split = (const float2 *)p;              p += sizeof(float2) * (size_t)Npad;
    size_t w32b = sizeof(float) * (size_t)N * K;
    w32b = (w32b + 127) / 128 * 128;
    v.w32 = (const float *)p;                    p += w32b;
    size_t w16b = sizeof(__half) * (size_t)Npad * K;
    w16b = (w16b + 127) / 128 * 128;
    size_t wtb = sizeof(__half) * (size_t)(Npad + 16 * YC_MAX_ANCHORS) * K;   // room for any per-anchor padding
    wtb = (wtb + 127) / 128 * 128;
    v.w_hi_t = (const __half *)p;                p += wtb;
    v.w_lo_t = (const __half *)p;                p += wtb;
    v.w_bf = (const __nv_bfloat16 *)p;
    return v;
}

// ---- numerics shared by every decode epilogue ---------------------------------------------
// sigmoid in binary32: 1 / (1 + exp(-t)), relative error ~2e-7, inside the 1e-5 parity bound with
// margin.  ONE special-function (MUFU) operation per value: the z-writing epilogues are bound by that pipe -- ncu on the
// IBin forward: XU pipe 67 % busy at 17 cycles per warp instruction, with ex2 + rcp per sigmoid -- so the reciprocal runs
// on the FMA pipe instead: exponent-trick seed (12 % off) and three Newton steps r <- r (2 - y r) (1.4e-2, 2e-4, 4e-8).
// ex2.approx.ftz: a flushed denormal cannot change 1 + e.  The clamp keeps y finite (t < -88: sigmoid -> 1e-38 instead of
// NaN from inf * 0).
__device__ __forceinline__ float sigmoidf_fast(float t)
{
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * -1.4426950408889634f));
    const float y = 1.0f + fminf(e, 1.0e38f);
    float r = __int_as_float(0x7EF311C7 - __float_as_int(y));
    r = r * fmaf(-y, r, 2.0f);
    r = r * fmaf(-y, r, 2.0f);
    r = r * fmaf(-y, r, 2.0f);
    return r;
}

// The two-MUFU form (ex2 + rcp, five instructions instead of thirteen) for epilogues that are bound by instruction issue
// rather than by the special-function pipe: the IBin half-row epilogues run two warps per scheduler (A/B on one box,
// 16 images at 1280x1280: forward 254.7 -> 227.3 us; the IDetect forward with three warps per scheduler loses 3.5 %).
// Every IBin path of one configuration uses the same form, so the fused step and the two-call path stay bit-identical.
__device__ __forceinline__ float sigmoidf_rcp(float t)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// xy = (s*2 - 0.5 + g) * stride, evaluated in the reference's operation order
// (nets/idetect.py:41) with no fused multiply-add.
__device__ __forceinline__ float decode_xy(float s, float g, float stride)
{
    float v = __fmul_rn(s, 2.0f);
    v = __fadd_rn(v, -0.5f);
    v = __fadd_rn(v, g);
    return __fmul_rn(v, stride);
}

// wh = (s*2)**2 * anchor (nets/idetect.py:42)
__device__ __forceinline__ float decode_wh(float s, float anchor)
{
    float v = __fmul_rn(s, 2.0f);
    v = __fmul_rn(v, v);
    return __fmul_rn(v, anchor);
}

} // namespace yc
