// Training-time augmentation on the device: what augmentation.py:16-234 does per example with a chain
// of TensorFlow image ops -- expand on a mean-filled canvas, crop + bilinear resize back to the input resolution,
// horizontal flip, brightness, contrast, hue, saturation, clip -- batched over the examples of a step.  Every random
// decision is an INPUT (one 16-word plan per image, drawn on the host in the reference's order), so the result can be
// compared with the reference source run on the same draws.  All kernels are HBM-bound gathers / streams: one read of
// the batch for the canvas mean (only images that expand), one read + one write for the geometric pass (brightness and,
// when no contrast follows, hue / saturation / clip fused in), one read + write for images that change contrast (the
// per-channel mean of the brightness-adjusted image is a grid-wide dependency).  Sums are combined in a fixed order:
// deterministic, no atomics.  float32, one rounding per operation like separate TensorFlow ops (-fmad=false).

#include "common.cuh"

namespace ssd {

struct AugPlan {                 // mirrors tf_ssd_b200/augmentation.py:pack_plans
    int flags;                   // bit0 patch, bit1 expand, bit2 flip, bit3 brightness, bit4 contrast, bit5 hue, bit6 saturation, bit7 no final clip
    int pad_top, pad_left, canvas_h, canvas_w;      // augmentation.py:177-184 (whole pixels)
    int crop_y0, crop_x0, crop_h, crop_w;           // sample_distorted_bounding_box window on the canvas
    float brightness, contrast, hue, saturation;    // drawn delta / factor of augmentation.py:67-116
    int reserved[3];
};
static_assert(sizeof(AugPlan) == 64, "plan is 16 words");

constexpr int AUG_PATCH = 1, AUG_EXPAND = 2, AUG_FLIP = 4, AUG_BRIGHT = 8, AUG_CONTRAST = 16, AUG_HUE = 32, AUG_SAT = 64,
              AUG_NOCLIP = 128;
constexpr int AUG_THREADS = 256;

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// documented tf.image.rgb_to_hsv / hsv_to_rgb formulas ([TF-recall]: TensorFlow's fused adjust_hue kernel agrees to
// float tolerance only)
__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
    v = fmaxf(fmaxf(r, g), b);
    const float spread = fsub(v, fminf(fminf(r, g), b));
    s = v > 0.0f ? fdiv(spread, v) : 0.0f;
    const float norm = fdiv(1.0f, fmul(6.0f, spread));
    float hh;
    if (r == v)      hh = fmul(norm, fsub(g, b));
    else if (g == v) hh = fadd(fmul(norm, fsub(b, r)), (float)(2.0 / 6.0));
    else             hh = fadd(fmul(norm, fsub(r, g)), (float)(4.0 / 6.0));
    hh = spread > 0.0f ? hh : 0.0f;
    h = hh < 0.0f ? fadd(hh, 1.0f) : hh;
}
__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
    const float dh = fmul(h, 6.0f);
    const float dr = clip01(fsub(fabsf(fsub(dh, 3.0f)), 1.0f));
    const float dg = clip01(fsub(2.0f, fabsf(fsub(dh, 2.0f))));
    const float db = clip01(fsub(2.0f, fabsf(fsub(dh, 4.0f))));
    const float oms = fsub(1.0f, s);
    r = fmul(fadd(oms, fmul(s, dr)), v);
    g = fmul(fadd(oms, fmul(s, dg)), v);
    b = fmul(fadd(oms, fmul(s, db)), v);
}

// hue (augmentation.py:93-104), saturation (:107-116), clip (:32)
__device__ __forceinline__ void photometric_tail(const AugPlan& p, float& r, float& g, float& b) {
    if (p.flags & AUG_HUE) {
        float h, s, v;
        rgb_to_hsv(r, g, b, h, s, v);
        h = fadd(h, p.hue);
        h = fsub(h, floorf(h));
        hsv_to_rgb(h, s, v, r, g, b);
    }
    if (p.flags & AUG_SAT) {
        float h, s, v;
        rgb_to_hsv(r, g, b, h, s, v);
        s = clip01(fmul(s, p.saturation));
        hsv_to_rgb(h, s, v, r, g, b);
    }
    if (!(p.flags & AUG_NOCLIP)) { r = clip01(r); g = clip01(g); b = clip01(b); }
}

// fixed-order CTA sum of three per-thread values -> partial[0..2]
__device__ __forceinline__ void block_sum3(float a, float b, float c, float* __restrict__ partial) {
    __shared__ float red[3][AUG_THREADS / 32];
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < AUG_THREADS / 32; ++w) s += red[threadIdx.x][w];
        partial[threadIdx.x] = s;
    }
}

// mean[c] = (sum of the image's per-CTA partial rows, in CTA order) / (H*W)
__device__ __forceinline__ void mean_from_partials(const float* __restrict__ partial, int nblk, int HW, float* __restrict__ mean_s) {
    if (threadIdx.x < 3) {
        float s = 0.0f;
        for (int k = 0; k < nblk; ++k) s += partial[k * 3 + threadIdx.x];
        mean_s[threadIdx.x] = fdiv(s, (float)HW);
    }
    __syncthreads();
}

// tf.nn.moments(img, [0, 1]) of augmentation.py:186, first stage: per-CTA channel sums of the images that expand
__global__ void __launch_bounds__(AUG_THREADS)
augment_sum_kernel(const float* __restrict__ img, int HW, const AugPlan* __restrict__ plans, float* __restrict__ partial) {
    const int b = blockIdx.y;
    if (!(plans[b].flags & AUG_EXPAND)) return;
    const float* src = img + (size_t)b * HW * 3;
    const int chunk = (HW + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * chunk, hi = min(HW, lo + chunk);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    for (int i = lo + threadIdx.x; i < hi; i += AUG_THREADS) {
        s0 += src[i * 3 + 0]; s1 += src[i * 3 + 1]; s2 += src[i * 3 + 2];
    }
    block_sum3(s0, s1, s2, partial + ((size_t)b * gridDim.x + blockIdx.x) * 3);
}

// canvas pixel (cy, cx) of augmentation.py:187-193: the image where it lies, its mean elsewhere
__device__ __forceinline__ void canvas_at(const float* __restrict__ src, const AugPlan& p, int H, int W, const float* mean_s,
                                          int cy, int cx, float& r, float& g, float& b) {
    const int y = cy - p.pad_top, x = cx - p.pad_left;
    if (y >= 0 && y < H && x >= 0 && x < W) {
        const float* q = src + ((size_t)y * W + x) * 3;
        r = q[0]; g = q[1]; b = q[2];
        if (p.flags & AUG_EXPAND) {                      // tf.where(expanded == -1, mean, expanded)
            if (r == -1.0f) r = mean_s[0];
            if (g == -1.0f) g = mean_s[1];
            if (b == -1.0f) b = mean_s[2];
        }
    } else {
        r = mean_s[0]; g = mean_s[1]; b = mean_s[2];
    }
}

// patch (augmentation.py:205-234) + flip (:119-139) + brightness (:67-78); hue / saturation / clip too when the
// image has no contrast step.  Gather form: one thread per output pixel.
__global__ void __launch_bounds__(AUG_THREADS)
augment_geometric_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W, int Ho, int Wo,
                         const AugPlan* __restrict__ plans, const float* __restrict__ in_partial, float* __restrict__ out_partial) {
    __shared__ float mean_s[3];
    __shared__ AugPlan plan_s;
    const int b = blockIdx.y, HW = H * W, HWo = Ho * Wo;
    if (threadIdx.x < 16) reinterpret_cast<int*>(&plan_s)[threadIdx.x] = reinterpret_cast<const int*>(plans + b)[threadIdx.x];
    __syncthreads();
    const AugPlan& p = plan_s;
    if (p.flags & AUG_EXPAND) mean_from_partials(in_partial + (size_t)b * gridDim.x * 3, gridDim.x, HW, mean_s);
    const float* src = img + (size_t)b * HW * 3;
    float* dst = out + (size_t)b * HWo * 3;
    const int chunk = (HWo + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * chunk, hi = min(HWo, lo + chunk);
    const bool patch = p.flags & AUG_PATCH, flip = p.flags & AUG_FLIP, contrast = p.flags & AUG_CONTRAST;
    const float scale_y = fdiv((float)p.crop_h, (float)Ho), scale_x = fdiv((float)p.crop_w, (float)Wo);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    for (int i = lo + threadIdx.x; i < hi; i += AUG_THREADS) {
        const int oy = i / Wo, ox_out = i - oy * Wo;
        const int ox = flip ? (Wo - 1 - ox_out) : ox_out;
        float r, g, bl;
        if (!patch) {
            const float* q = src + ((size_t)oy * W + ox) * 3;
            r = q[0]; g = q[1]; bl = q[2];
        } else {
            const float in_y = fsub(fmul(fadd((float)oy, 0.5f), scale_y), 0.5f);
            const float in_x = fsub(fmul(fadd((float)ox, 0.5f), scale_x), 0.5f);
            const float fy = floorf(in_y), fx = floorf(in_x);
            const int y0 = max((int)fy, 0) + p.crop_y0, y1 = min((int)ceilf(in_y), p.crop_h - 1) + p.crop_y0;
            const int x0 = max((int)fx, 0) + p.crop_x0, x1 = min((int)ceilf(in_x), p.crop_w - 1) + p.crop_x0;
            const float ly = fsub(in_y, fy), lx = fsub(in_x, fx);
            float tl[3], tr[3], b0[3], b1[3];
            canvas_at(src, p, H, W, mean_s, y0, x0, tl[0], tl[1], tl[2]);
            canvas_at(src, p, H, W, mean_s, y0, x1, tr[0], tr[1], tr[2]);
            canvas_at(src, p, H, W, mean_s, y1, x0, b0[0], b0[1], b0[2]);
            canvas_at(src, p, H, W, mean_s, y1, x1, b1[0], b1[1], b1[2]);
            float v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float top = fadd(tl[c], fmul(fsub(tr[c], tl[c]), lx));
                const float bot = fadd(b0[c], fmul(fsub(b1[c], b0[c]), lx));
                v[c] = fadd(top, fmul(fsub(bot, top), ly));
            }
            r = v[0]; g = v[1]; bl = v[2];
        }
        if (p.flags & AUG_BRIGHT) { r = fadd(r, p.brightness); g = fadd(g, p.brightness); bl = fadd(bl, p.brightness); }
        if (contrast) { s0 += r; s1 += g; s2 += bl; }
        else photometric_tail(p, r, g, bl);
        dst[(size_t)i * 3 + 0] = r; dst[(size_t)i * 3 + 1] = g; dst[(size_t)i * 3 + 2] = bl;
    }
    if (contrast) block_sum3(s0, s1, s2, out_partial + ((size_t)b * gridDim.x + blockIdx.x) * 3);
}

// contrast (augmentation.py:81-90: (x - mean) * factor + mean, per-channel mean over height and width) followed by
// hue / saturation / clip, in place, for the images that have a contrast step
__global__ void __launch_bounds__(AUG_THREADS)
augment_contrast_kernel(float* __restrict__ out, int HW, const AugPlan* __restrict__ plans, const float* __restrict__ out_partial) {
    __shared__ float mean_s[3];
    __shared__ AugPlan plan_s;
    const int b = blockIdx.y;
    if (!(plans[b].flags & AUG_CONTRAST)) return;
    if (threadIdx.x < 16) reinterpret_cast<int*>(&plan_s)[threadIdx.x] = reinterpret_cast<const int*>(plans + b)[threadIdx.x];
    __syncthreads();
    const AugPlan& p = plan_s;
    mean_from_partials(out_partial + (size_t)b * gridDim.x * 3, gridDim.x, HW, mean_s);
    float* dst = out + (size_t)b * HW * 3;
    const int chunk = (HW + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * chunk, hi = min(HW, lo + chunk);
    for (int i = lo + threadIdx.x; i < hi; i += AUG_THREADS) {
        float r = dst[(size_t)i * 3 + 0], g = dst[(size_t)i * 3 + 1], bl = dst[(size_t)i * 3 + 2];
        r = fadd(fmul(fsub(r, mean_s[0]), p.contrast), mean_s[0]);
        g = fadd(fmul(fsub(g, mean_s[1]), p.contrast), mean_s[1]);
        bl = fadd(fmul(fsub(bl, mean_s[2]), p.contrast), mean_s[2]);
        photometric_tail(p, r, g, bl);
        dst[(size_t)i * 3 + 0] = r; dst[(size_t)i * 3 + 1] = g; dst[(size_t)i * 3 + 2] = bl;
    }
}

// boxes: expand_image's and patch's renormalize_bboxes_with_min_max (utils/bbox_utils.py:217-233), then the flip
// (augmentation.py:128-137).  All-zero rows are batch padding (utils/data_utils.py:140-155) and stay zero.
__device__ __forceinline__ float4 renorm(float4 bx, float y_min, float x_min, float y_max, float x_max) {
    const float sy = fsub(y_max, y_min), sx = fsub(x_max, x_min);
    return make_float4(clip01(fdiv(fsub(bx.x, y_min), sy)), clip01(fdiv(fsub(bx.y, x_min), sx)),
                       clip01(fdiv(fsub(bx.z, y_min), sy)), clip01(fdiv(fsub(bx.w, x_min), sx)));
}
__global__ void augment_boxes_kernel(float4* __restrict__ boxes, int G, int total, const AugPlan* __restrict__ plans, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float4 bx = boxes[i];
    if (bx.x == 0.f && bx.y == 0.f && bx.z == 0.f && bx.w == 0.f) return;
    const AugPlan p = plans[i / G];
    if (p.flags & AUG_PATCH) {
        if (p.flags & AUG_EXPAND) {
            const float h = (float)H, w = (float)W;
            const float pad_bottom = (float)(p.canvas_h - (H + p.pad_top)), pad_right = (float)(p.canvas_w - (W + p.pad_left));
            bx = renorm(bx, fdiv(-(float)p.pad_top, h), fdiv(-(float)p.pad_left, w), fdiv(fadd(pad_bottom, h), h),
                        fdiv(fadd(pad_right, w), w));
        }
        const float ch = (float)p.canvas_h, cw = (float)p.canvas_w;
        bx = renorm(bx, fdiv((float)p.crop_y0, ch), fdiv((float)p.crop_x0, cw), fdiv((float)(p.crop_y0 + p.crop_h), ch),
                    fdiv((float)(p.crop_x0 + p.crop_w), cw));
    }
    if (p.flags & AUG_FLIP) bx = make_float4(bx.x, fsub(1.0f, bx.w), bx.z, fsub(1.0f, bx.y));
    boxes[i] = bx;
}

static int aug_blocks_per_image(int B, int HW) {
    // enough CTAs for ~4 waves of the machine, at least 1024 pixels each
    const int want = max(1, (sm_count() * 8 + B - 1) / B);
    return max(1, min(want, (HW + 1023) / 1024));
}

}  // namespace ssd

using namespace ssd;

extern "C" size_t ssd_augment_workspace_bytes(int B, int H, int W, int out_h, int out_w) {
    if (B < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return 0;
    return (size_t)2 * B * aug_blocks_per_image(B, max(H * W, out_h * out_w)) * 3 * sizeof(float);
}

extern "C" int ssd_augment_batch(const float* d_images, float* d_out, float* d_boxes, int B, int H, int W, int out_h,
                                 int out_w, int G, const void* d_plans, void* d_workspace, size_t workspace_bytes, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_images); SSD_REQUIRE_PTR(d_out); SSD_REQUIRE_PTR(d_plans); SSD_REQUIRE_PTR(d_workspace);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && out_h >= 1 && out_w >= 1 && G >= 0 && (int64_t)H * W < (1 << 28) &&
                (int64_t)out_h * out_w < (1 << 28) && B <= 65535, SSD_ERR_SHAPE,
                "ssd_augment_batch: bad shape B=%d H=%d W=%d out=%dx%d G=%d", B, H, W, out_h, out_w, G);
    SSD_REQUIRE(d_images != d_out, SSD_ERR_SHAPE, "ssd_augment_batch: the geometric pass is a gather, d_out must not alias d_images");
    SSD_REQUIRE(G == 0 || d_boxes != nullptr, SSD_ERR_NULL, "ssd_augment_batch: d_boxes is NULL");
    const size_t need = ssd_augment_workspace_bytes(B, H, W, out_h, out_w);
    SSD_REQUIRE(workspace_bytes >= need, SSD_ERR_WORKSPACE, "ssd_augment_batch: workspace %zu < %zu bytes", workspace_bytes, need);
    const int HW = H * W, HWo = out_h * out_w, nblk = aug_blocks_per_image(B, max(HW, HWo));
    const AugPlan* plans = static_cast<const AugPlan*>(d_plans);
    float* in_partial = static_cast<float*>(d_workspace);
    float* out_partial = in_partial + (size_t)B * nblk * 3;
    cudaStream_t st = as_stream(stream);
    const dim3 grid(nblk, B);
    augment_sum_kernel<<<grid, AUG_THREADS, 0, st>>>(d_images, HW, plans, in_partial);
    SSD_CHECK_LAUNCH("augment_sum_kernel");
    augment_geometric_kernel<<<grid, AUG_THREADS, 0, st>>>(d_images, d_out, H, W, out_h, out_w, plans, in_partial, out_partial);
    SSD_CHECK_LAUNCH("augment_geometric_kernel");
    augment_contrast_kernel<<<grid, AUG_THREADS, 0, st>>>(d_out, HWo, plans, out_partial);
    SSD_CHECK_LAUNCH("augment_contrast_kernel");
    if (G > 0) {
        augment_boxes_kernel<<<ceil_div((int64_t)B * G, 128), 128, 0, st>>>(reinterpret_cast<float4*>(d_boxes), G, B * G, plans, H, W);
        SSD_CHECK_LAUNCH("augment_boxes_kernel");
    }
    return SSD_OK;
}
