// yc_preproc.cu -- the two steps either side of the model that SURVEY.md section 8(f) ranks next:
//   * letterbox preprocessing of a batch of decoded images (reference detect.py:16-26 prepare_test_image +
//     image_enhance/letter_box.py:27-60): bilinear resize, 114-pad, /255, HWC -> CHW in ONE pass;
//   * formatting of the detections (reference detect.py:236-258): label, confidence, floored and clamped boxes.
#include "yc_common.cuh"

namespace yc {

// cv2.resize(..., INTER_LINEAR) for 8-bit images is fixed point (third-party OpenCV, resize.cpp: 11-bit
// coefficients, horizontal pass into int32, vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2).
// The coefficient of output coordinate d: f = (float)((d + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in
// binary64; s = floor(f); f -= s.  Horizontally s is clamped to [0, src-1] with f = 0 at the borders; vertically
// only the ROW INDICES are clamped (both rows can be the same row, each product truncated separately).
// Pinned bit for bit against cv2 4.13 (tests/golden/letterbox_*.npz).
struct LinCoef { int s; int a0, a1; };

__device__ __forceinline__ LinCoef lin_coef(int d, double scale, int ssize, bool clamp)
{
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    LinCoef c;
    c.s = s;
    c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    c.a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    return c;
}

template <typename OUT> __device__ __forceinline__ OUT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// grid (ceil(out_w / 128), out_h, bs), 128 threads: one output pixel (3 channels) per thread; the three planes are
// written with consecutive threads on consecutive x (coalesced), the source is read through L1/L2 (4 pixels per
// output pixel, neighbouring threads share them).
template <typename OUT>
__global__ void __launch_bounds__(128) letterbox_kernel(const yc_letterbox_image *__restrict__ imgs, int out_h, int out_w,
                                                        OUT *__restrict__ out)
{
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= out_w) return;
    const yc_letterbox_image im = imgs[b];
    float v[3];
    const int rx = x - im.left, ry = y - im.top;
    if (rx < 0 || rx >= im.rs_w || ry < 0 || ry >= im.rs_h) {
        v[0] = v[1] = v[2] = (float)im.pad_value;
    } else {
        const double sx = 1.0 / ((double)im.rs_w / (double)im.src_w), sy = 1.0 / ((double)im.rs_h / (double)im.src_h);
        const LinCoef cx = lin_coef(rx, sx, im.src_w, true), cy = lin_coef(ry, sy, im.src_h, false);
        const int x0 = cx.s, x1 = min(cx.s + 1, im.src_w - 1);
        const int y0 = min(max(cy.s, 0), im.src_h - 1), y1 = min(max(cy.s + 1, 0), im.src_h - 1);
        const uint8_t *r0 = im.src + (size_t)y0 * im.src_pitch, *r1 = im.src + (size_t)y1 * im.src_pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int s0 = (int)r0[3 * x0 + c] * cx.a0 + (int)r0[3 * x1 + c] * cx.a1;
            const int s1 = (int)r1[3 * x0 + c] * cx.a0 + (int)r1[3 * x1 + c] * cx.a1;
            const int q = (((cy.a0 * (s0 >> 4)) >> 16) + ((cy.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
            v[c] = (float)min(max(q, 0), 255);
        }
    }
    const size_t plane = (size_t)out_h * out_w;
    OUT *o = out + (size_t)b * 3 * plane + (size_t)y * out_w + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = to_out<OUT>(__fdiv_rn(v[c], 255.0f)); // np.float32(image) / 255.
}

// detect.py:236-258 -- rows hold (y1, x1, y2, x2, obj, class_conf, class) in image pixels
__global__ void __launch_bounds__(256) format_kernel(const float *__restrict__ rows, const int32_t *__restrict__ offsets,
                                                     const int32_t *__restrict__ image_hw, int image_hw_stride,
                                                     int32_t *__restrict__ box_xyxy, float *__restrict__ conf,
                                                     int32_t *__restrict__ label)
{
    const int b = blockIdx.y;
    const int lo = offsets[b], hi = offsets[b + 1];
    const int ih = image_hw[b * image_hw_stride], iw = image_hw[b * image_hw_stride + 1];
    for (int i = lo + blockIdx.x * 256 + threadIdx.x; i < hi; i += gridDim.x * 256) {
        const float *r = rows + (size_t)i * 7;
        label[i] = (int32_t)r[6];                        // np.array(results[:, 6], dtype='int32')
        conf[i] = __fmul_rn(r[4], r[5]);                 // results[:, 4] * results[:, 5]
        int4 o;
        o.x = max(0, (int)floorf(r[1]));                 // x1 = max(0, floor(x1))
        o.y = max(0, (int)floorf(r[0]));
        o.z = min(iw, (int)floorf(r[3]));                // x2 = min(image width, floor(x2))
        o.w = min(ih, (int)floorf(r[2]));
        ((int4 *)box_xyxy)[i] = o;
    }
}

} // namespace yc

using namespace yc;

extern "C" int yc_letterbox_batch(const yc_letterbox_image *imgs, int bs, int out_h, int out_w, int out_dtype, void *out,
                                  yc_stream_t stream)
{
    YC_REQUIRE(imgs && out, YC_ERR_INVALID, "yc_letterbox_batch: null argument");
    YC_REQUIRE(bs > 0 && bs <= 65535 && out_h > 0 && out_h <= 65535 && out_w > 0, YC_ERR_INVALID,
               "yc_letterbox_batch: bad shape (bs=%d, out=%dx%d)", bs, out_h, out_w);
    const dim3 grid((out_w + 127) / 128, out_h, bs);
    if (out_dtype == YC_F32)
        letterbox_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(imgs, out_h, out_w, (float *)out);
    else if (out_dtype == YC_BF16)
        letterbox_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(imgs, out_h, out_w, (__nv_bfloat16 *)out);
    else
        YC_REQUIRE(false, YC_ERR_INVALID, "yc_letterbox_batch: bad output dtype %d", out_dtype);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

extern "C" int yc_format_detections(const float *rows, const int32_t *offsets, int bs, const int32_t *image_hw,
                                    int image_hw_stride, int32_t *box_xyxy, float *conf, int32_t *label, yc_stream_t stream)
{
    YC_REQUIRE(rows && offsets && image_hw && box_xyxy && conf && label, YC_ERR_INVALID, "yc_format_detections: null argument");
    YC_REQUIRE(bs > 0 && bs <= 65535, YC_ERR_INVALID, "yc_format_detections: bad batch size %d", bs);
    YC_REQUIRE(((uintptr_t)box_xyxy & 15) == 0, YC_ERR_INVALID, "yc_format_detections: box_xyxy must be 16-byte aligned");
    format_kernel<<<dim3(4, bs), 256, 0, (cudaStream_t)stream>>>(rows, offsets, image_hw, image_hw_stride, box_xyxy, conf,
                                                                label);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}
