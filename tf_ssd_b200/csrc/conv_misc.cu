// HBM-bound layers of the two SSD graphs for sm_100a (NHWC fp16, 16-byte
// vector accesses, one pass over the data):
//   * DepthwiseConv2D 3x3 (+ folded BN + ReLU6)   -- MobileNetV2 blocks
//   * fp32 image -> fp16 with the channel axis padded to 8
//   * MaxPool2D(padding="same")                    -- VGG16 pool1..pool5
//   * L2Normalization                              -- VGG16 conv4_3 tap
// None of these is a dense contraction; they are not reshaped into GEMMs.

#include "common.cuh"

namespace ssd {

__device__ __forceinline__ void h8_to_f(const uint4 v, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 t = __half22float2(h[k]);
        f[2 * k] = t.x; f[2 * k + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 f_to_h8(const float (&f)[8]) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
    return o;
}

// ------------------------------------------------------------- depthwise --
// Each thread owns 8 channels of PX consecutive output pixels of one row, so
// the input columns shared by neighbouring outputs are loaded once.  Channel
// groups are the fastest thread index: a warp reads 512 contiguous bytes.
template <int STRIDE, int PX>
__global__ void __launch_bounds__(256)
depthwise3x3_kernel(const uint4* __restrict__ in, const uint4* __restrict__ w, const float* __restrict__ bias,
                    uint4* __restrict__ out, int H, int W, int C8, int Ho, int Wo, int pad_t, int pad_l,
                    int act, int64_t total) {
    constexpr int COLS = (PX - 1) * STRIDE + 3;
    pdl_trigger();
    pdl_wait();
    // grid = (chunks of one output row's (pixel group, channel group) items, Ho, B): one 32-bit division per
    // thread instead of the 64-bit div/mod chain of a flat index (that chain cost more issue slots than the
    // convolution itself)
    const int wgroups = (Wo + PX - 1) / PX;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    (void)total;
    if (item < wgroups * C8) {
        const int xg = item / C8;
        const int c8 = item - xg * C8;
        const int oy = blockIdx.y;
        const int b = blockIdx.z;
        const int ox0 = xg * PX;

        float acc[PX][8];
        float bv[8];
        if (bias) {
            float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + c8 * 2);
            float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + c8 * 2 + 1);
            bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
            bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) bv[k] = 0.0f;
        }
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[p][k] = bv[k];

        // All input vectors of the 3 x COLS window are requested before the first one is consumed (predicated
        // loads, zero for padding): 12 (stride 1) / 15 (stride 2) independent 16-byte loads in flight per thread --
        // this kernel is bound by memory latency x bytes in flight, not by arithmetic.
        const uint4* img = in + (size_t)b * H * W * C8;
        uint4 xin[3][COLS];
        uint4 wv[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * STRIDE - pad_t + ky;
            const bool vy = (unsigned)iy < (unsigned)H;
            const uint4* row = img + (size_t)(vy ? iy : 0) * W * C8 + c8;
#pragma unroll
            for (int cx = 0; cx < COLS; ++cx) {
                const int ix = ox0 * STRIDE - pad_l + cx;
                const bool v = vy && (unsigned)ix < (unsigned)W;
                xin[ky][cx] = v ? __ldg(row + (size_t)ix * C8) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) wv[ky][kx] = __ldg(w + (size_t)(ky * 3 + kx) * C8 + c8);
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float wk[3][8];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) h8_to_f(wv[ky][kx], wk[kx]);
#pragma unroll
            for (int cx = 0; cx < COLS; ++cx) {
                float x[8];
                h8_to_f(xin[ky][cx], x);
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const int kx = cx - p * STRIDE;
                    if (kx >= 0 && kx < 3) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[p][k] = fmaf(x[k], wk[kx][k], acc[p][k]);
                    }
                }
            }
        }
        uint4* orow = out + ((size_t)(b * Ho + oy) * Wo) * C8 + c8;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            if (ox0 + p >= Wo) break;
            if (act == SSD_ACT_RELU6) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[p][k] = fminf(fmaxf(acc[p][k], 0.0f), 6.0f);
            } else if (act == SSD_ACT_RELU) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[p][k] = fmaxf(acc[p][k], 0.0f);
            }
            orow[(size_t)(ox0 + p) * C8] = f_to_h8(acc[p]);
        }
    }
}

// ------------------------------------------------------- image -> fp16 c8 --
__global__ void __launch_bounds__(256)
image_to_f16c8_kernel(const float* __restrict__ img, uint4* __restrict__ out, int64_t n_pixels) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels;
         i += (int64_t)gridDim.x * blockDim.x) {
        float f[8] = {__ldcs(img + i * 3), __ldcs(img + i * 3 + 1), __ldcs(img + i * 3 + 2), 0.f, 0.f, 0.f, 0.f, 0.f};
        out[i] = f_to_h8(f);
    }
}

// ---------------------------------------------------------------- maxpool --
__global__ void __launch_bounds__(256)
maxpool_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8, int Ho, int Wo,
               int k, int stride, int pad_t, int pad_l, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int c8 = (int)(t % C8);
        int64_t r = t / C8;
        int ox = (int)(r % Wo); r /= Wo;
        int oy = (int)(r % Ho);
        int b = (int)(r / Ho);
        const __half2 ninf = __float2half2_rn(-INFINITY);
        __half2 m[4] = {ninf, ninf, ninf, ninf};
        for (int ky = 0; ky < k; ++ky) {
            int iy = oy * stride - pad_t + ky;
            if ((unsigned)iy >= (unsigned)H) continue;
            for (int kx = 0; kx < k; ++kx) {
                int ix = ox * stride - pad_l + kx;
                if ((unsigned)ix >= (unsigned)W) continue;
                uint4 v = __ldg(in + ((size_t)(b * H + iy) * W + ix) * C8 + c8);
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], h[q]);
            }
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) oh[q] = m[q];
        out[t] = o;
    }
}

// ----------------------------------------------------------------- l2norm --
// One warp per pixel row of C channels (C % 8 == 0).
__global__ void __launch_bounds__(256)
l2norm_kernel(const uint4* __restrict__ in, const float* __restrict__ scale, uint4* __restrict__ out,
              int64_t rows, int C8) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const uint4* src = in + r * C8;
        float ss = 0.0f;
        for (int c = lane; c < C8; c += 32) {
            float f[8];
            h8_to_f(__ldg(src + c), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) ss = fmaf(f[k], f[k], ss);
        }
        ss = warp_sum(ss);
        const float inv = rsqrtf(fmaxf(ss, 1e-12f));
        for (int c = lane; c < C8; c += 32) {
            float f[8];
            h8_to_f(__ldg(src + c), f);
            float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + c * 2);
            float4 s1 = __ldg(reinterpret_cast<const float4*>(scale) + c * 2 + 1);
            f[0] *= inv * s0.x; f[1] *= inv * s0.y; f[2] *= inv * s0.z; f[3] *= inv * s0.w;
            f[4] *= inv * s1.x; f[5] *= inv * s1.y; f[6] *= inv * s1.z; f[7] *= inv * s1.w;
            out[r * C8 + c] = f_to_h8(f);
        }
    }
}

static int grid_for(int64_t threads, int per_sm = 8) {
    int64_t blocks = (threads + 255) / 256;
    int64_t cap = (int64_t)sm_count() * per_sm;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_depthwise3x3(const void* d_in, const void* d_weight, const float* d_bias, void* d_out,
                                int B, int H, int W, int C, int Ho, int Wo, int stride, int pad_top, int pad_left,
                                int act, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_in); SSD_REQUIRE_PTR(d_weight); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && Ho >= 1 && Wo >= 1 &&
                (stride == 1 || stride == 2) && act >= SSD_ACT_NONE && act <= SSD_ACT_RELU6, SSD_ERR_SHAPE,
                "ssd_depthwise3x3: bad shape B=%d H=%d W=%d C=%d Ho=%d Wo=%d stride=%d act=%d (C %% 8 == 0, stride 1|2)",
                B, H, W, C, Ho, Wo, stride, act);
    const int C8 = C / 8;
    constexpr int PX = 2;
    const int64_t total = (int64_t)B * Ho * ((Wo + PX - 1) / PX) * C8;
    SSD_REQUIRE(Ho <= 65535 && B <= 65535, SSD_ERR_SHAPE, "ssd_depthwise3x3: Ho=%d / B=%d exceed the grid limits", Ho, B);
    const int row_items = ((Wo + PX - 1) / PX) * C8;
    auto args = [&](auto kern) {
        return launch_pdl(kern, dim3((row_items + 255) / 256, Ho, B), dim3(256), 0, as_stream(stream),
                          reinterpret_cast<const uint4*>(d_in), reinterpret_cast<const uint4*>(d_weight), d_bias,
                          reinterpret_cast<uint4*>(d_out), H, W, C8, Ho, Wo, pad_top, pad_left, act, total);
    };
    cudaError_t le = (stride == 1) ? args(depthwise3x3_kernel<1, PX>) : args(depthwise3x3_kernel<2, PX>);
    if (le != cudaSuccess) return cuda_fail(le, "depthwise3x3_kernel");
    return SSD_OK;
}

extern "C" int ssd_image_to_f16c8(const float* d_img, void* d_out, int64_t n_pixels, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_img); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(n_pixels >= 0, SSD_ERR_SHAPE, "ssd_image_to_f16c8: n_pixels=%lld", (long long)n_pixels);
    if (n_pixels == 0) return SSD_OK;
    image_to_f16c8_kernel<<<grid_for(n_pixels, 16), 256, 0, as_stream(stream)>>>(
        d_img, reinterpret_cast<uint4*>(d_out), n_pixels);
    SSD_CHECK_LAUNCH("image_to_f16c8_kernel");
    return SSD_OK;
}

extern "C" int ssd_maxpool(const void* d_in, void* d_out, int B, int H, int W, int C, int Ho, int Wo,
                           int k, int stride, int pad_top, int pad_left, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_in); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && Ho >= 1 && Wo >= 1 && k >= 1 && k <= 7 &&
                stride >= 1, SSD_ERR_SHAPE, "ssd_maxpool: bad shape B=%d H=%d W=%d C=%d Ho=%d Wo=%d k=%d s=%d",
                B, H, W, C, Ho, Wo, k, stride);
    const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
    maxpool_kernel<<<grid_for(total, 16), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_out), H, W, C / 8, Ho, Wo, k, stride,
        pad_top, pad_left, total);
    SSD_CHECK_LAUNCH("maxpool_kernel");
    return SSD_OK;
}

extern "C" int ssd_l2norm(const void* d_in, const float* d_scale, void* d_out, int64_t rows, int C,
                          ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_in); SSD_REQUIRE_PTR(d_scale); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(rows >= 0 && C >= 8 && C % 8 == 0, SSD_ERR_SHAPE, "ssd_l2norm: bad shape rows=%lld C=%d",
                (long long)rows, C);
    if (rows == 0) return SSD_OK;
    l2norm_kernel<<<grid_for(rows * 32, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint4*>(d_in), d_scale, reinterpret_cast<uint4*>(d_out), rows, C / 8);
    SSD_CHECK_LAUNCH("l2norm_kernel");
    return SSD_OK;
}
