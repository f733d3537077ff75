// Device-side input pipeline: what utils/data_utils.py:33-37 does per example on the host with
// TensorFlow -- tf.image.convert_image_dtype(uint8 -> float32) followed by tf.image.resize (bilinear, half-pixel
// centres) -- plus augmentation.py:flip_horizontally (:119-139), fused into one HBM-bound pass that writes the
// example straight into its slot of the NHWC float32 batch the network reads.

#include "common.cuh"

namespace ssd {

// [TF-recall] tensorflow/core/kernels/image/resize_bilinear_op.cc with half_pixel_centers = true:
//   in = (out + 0.5) * scale - 0.5;  lower = max(floor(in), 0);  upper = min(ceil(in), size - 1);  lerp = in - floor(in)
//   value = top + (bottom - top) * y_lerp,  top = tl + (tr - tl) * x_lerp   (float32, one rounding per operation)
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ img, int H, int W, float* __restrict__ out, int S_h, int S_w,
                  float scale_y, float scale_x, int flip) {
    const int total = S_h * S_w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int oy = i / S_w, ox_out = i - oy * S_w;
        const int ox = flip ? (S_w - 1 - ox_out) : ox_out;               // flip_left_right of the RESIZED image
        const float in_y = fsub(fmul(fadd((float)oy, 0.5f), scale_y), 0.5f);
        const float in_x = fsub(fmul(fadd((float)ox, 0.5f), scale_x), 0.5f);
        const float fy = floorf(in_y), fx = floorf(in_x);
        const int y0 = max((int)fy, 0), y1 = min((int)ceilf(in_y), H - 1);
        const int x0 = max((int)fx, 0), x1 = min((int)ceilf(in_x), W - 1);
        const float ly = fsub(in_y, fy), lx = fsub(in_x, fx);
        const uint8_t* r0 = img + (size_t)y0 * W * 3;
        const uint8_t* r1 = img + (size_t)y1 * W * 3;
        const float k = 1.0f / 255.0f;                                   // convert_image_dtype: cast * float32(1/255)
        float* dst = out + (size_t)i * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float tl = fmul((float)r0[x0 * 3 + c], k), tr = fmul((float)r0[x1 * 3 + c], k);
            const float bl = fmul((float)r1[x0 * 3 + c], k), br = fmul((float)r1[x1 * 3 + c], k);
            const float top = fadd(tl, fmul(fsub(tr, tl), lx));
            const float bot = fadd(bl, fmul(fsub(br, bl), lx));
            dst[c] = fadd(top, fmul(fsub(bot, top), ly));
        }
    }
}

// augmentation.py:128-137: [y1, 1 - x2, y2, 1 - x1]; padded (all-zero) boxes stay zero
__global__ void flip_boxes_kernel(float4* __restrict__ boxes, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = boxes[i];
    if (b.x == 0.f && b.y == 0.f && b.z == 0.f && b.w == 0.f) return;
    boxes[i] = make_float4(b.x, fsub(1.0f, b.w), b.z, fsub(1.0f, b.y));
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_preprocess_image(const void* d_img_u8, int H, int W, float* d_out, int out_h, int out_w, int flip,
                                    ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_img_u8); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(H >= 1 && W >= 1 && out_h >= 1 && out_w >= 1 && (int64_t)out_h * out_w < (1 << 30), SSD_ERR_SHAPE,
                "ssd_preprocess_image: bad shape H=%d W=%d out=%dx%d", H, W, out_h, out_w);
    const int total = out_h * out_w;
    const int blocks = min(ceil_div(total, 256), sm_count() * 8);
    preprocess_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(d_img_u8), H, W, d_out, out_h, out_w,
                                                             (float)H / (float)out_h, (float)W / (float)out_w, flip ? 1 : 0);
    SSD_CHECK_LAUNCH("preprocess_kernel");
    return SSD_OK;
}

extern "C" int ssd_flip_boxes(float* d_boxes, int n, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_boxes);
    SSD_REQUIRE(n >= 0, SSD_ERR_SHAPE, "ssd_flip_boxes: n=%d", n);
    if (n == 0) return SSD_OK;
    flip_boxes_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(d_boxes), n);
    SSD_CHECK_LAUNCH("flip_boxes_kernel");
    return SSD_OK;
}
