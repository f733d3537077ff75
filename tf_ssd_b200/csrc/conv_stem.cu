// MobileNetV2 stem: Conv2D 3 -> 32, 3x3, stride 2 (keras_applications Conv1 + folded bn_Conv1 + ReLU6), read straight from
// the NHWC image -- float32 in [0,1] (what utils/data_utils.py:36 hands the reference model) or uint8 (what the
// reference's pipeline holds BEFORE tf.image.convert_image_dtype, data_utils.py:33-37; the conversion
// float32(u8) * float32(1/255) is fused here, so a uint8 batch costs a quarter of the H2D and HBM bytes).
//
// K = 27 is too shallow for the tcgen05 path (one 128 x 32 x 32 tile per 128 pixels would spend its time in barrier
// and TMEM hand-offs), and the scalar version spent 864 FMAs per output pixel on the CUDA cores (64 us per batch of
// 32, five times the HBM time).  Here one CTA owns one output row:
//   1. the three input rows it needs are staged in shared memory as fp16 with coalesced 16-byte (f32) / 4-byte (u8)
//      global loads -- a 3x3x3 window is then three runs of 9 contiguous halves;
//   2. each warp multiplies 32 output pixels by the [27(+5 zero) x 32] filter with mma.sync.m16n8k16 (fp16 operands,
//      fp32 accumulate): the A fragments are gathered directly from the staged rows, B lives in registers;
//   3. bias + activation, fp16 results through a padded shared tile, 16-byte coalesced stores (an output row is one
//      contiguous run of Wo * 64 bytes).
// HBM-bound by design: reads the image once (+ one row in three re-read through L2), writes the activation once.

#include "common.cuh"

namespace ssd {

constexpr int STEM_THREADS = 160;                  // 5 warps x 32 output pixels = one chunk of 160 pixels of a row
constexpr int STEM_PIX = STEM_THREADS;
constexpr int STEM_ROWLEN = STEM_PIX * 6 + 16;     // halves per staged row (pixel p's window starts at 3 * stride * p)
// halves per pixel in the output tile: Cout + 8 (80 B / 144 B: conflict-free, 16-byte aligned)

__device__ __forceinline__ float stem_to_float(float v) { return v; }
__device__ __forceinline__ float stem_to_float(uint8_t v) { return __fmul_rn((float)v, 1.0f / 255.0f); }
// third input form: the fp16 NHWC image with the channel axis padded to 8 (what ssd_image_to_f16c8 writes and what the
// first layer's filter gradient reads in the training plan)
struct HalfC8 { __half c[8]; };

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// STRIDE 2, NT = 4: MobileNetV2's Conv1 (3 -> 32).  STRIDE 1, NT = 8: VGG16's conv1_1 (3 -> 64, models/ssd_vgg16.py:80) --
// as a tensor-core layer its K = 27 would be padded to 9 k-blocks of 64 channels (1.1 ms per batch of 32, a third of the
// whole VGG16 forward); here it is bound by writing its 369 MB output.
template <typename TIn, int STRIDE, int NT>
__global__ void __launch_bounds__(STEM_THREADS)
stem_conv3x3s2_mma_kernel(const TIn* __restrict__ img, const __half* __restrict__ w, const float* __restrict__ bias,
                          __half* __restrict__ out, int H, int W, int Ho, int Wo, int pad_t, int pad_l, int act, int chunks) {
    constexpr int WCIN = sizeof(TIn) == 16 ? 8 : 3;           // input channels per tap in the weight layout
    constexpr int STEP = 3 * STRIDE;                          // staged halves between neighbouring output pixels
    constexpr int COUT = 8 * NT, STEM_OSTRIDE = COUT + 8;
    __shared__ __align__(16) __half srow[3][STEM_ROWLEN];
    __shared__ __align__(16) __half sout[STEM_PIX * STEM_OSTRIDE];
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int oy = blockIdx.x / chunks, chunk = blockIdx.x - oy * chunks, b = blockIdx.y;
    const int ox0 = chunk * STEM_PIX;                       // first output pixel of this CTA
    const int ix_first = ox0 * STRIDE - pad_l;              // input column of staged element 0

    // B fragments (weights OHWI [Cout][27] fp16, k = (ky*3+kx)*3+ci, zero for k >= 27): n = 8 j + g
    uint32_t bf[2][NT][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                // weights OHWI [Cout][3][3][w_cin] (w_cin = 3, or 8 for the zero-padded layout of the tensor-map path)
                const int k = 16 * s + 8 * r + 2 * t;
                const __half* wr = w + (8 * j + g) * 9 * WCIN;
                const int i0 = WCIN == 3 ? k : (k / 3) * WCIN + k % 3, i1 = WCIN == 3 ? k + 1 : ((k + 1) / 3) * WCIN + (k + 1) % 3;
                const __half lo = k < 27 ? wr[i0] : __float2half(0.f);
                const __half hi = k + 1 < 27 ? wr[i1] : __float2half(0.f);
                bf[s][j][r] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
    pdl_wait();

    // ---- stage the three input rows as fp16 (zero outside the image) -------------------------------------
    const int need = min(STEM_PIX, Wo - ox0) * STEP + (9 - STEP);     // staged halves actually read by valid pixels
    const TIn* base = img + (size_t)b * H * W * 3;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * STRIDE - pad_t + ky;
        __half* dst = srow[ky];
        if ((unsigned)iy >= (unsigned)H) {
            for (int e = tid; e < STEM_ROWLEN; e += STEM_THREADS) dst[e] = __float2half(0.f);
            continue;
        }
        const int e_lo = max(0, -ix_first * 3), e_hi = min(need, (W - ix_first) * 3);     // valid range [e_lo, e_hi)
        for (int e = tid; e < e_lo; e += STEM_THREADS) dst[e] = __float2half(0.f);
        for (int e = e_hi + tid; e < STEM_ROWLEN; e += STEM_THREADS) dst[e] = __float2half(0.f);
        if constexpr (sizeof(TIn) == 16) {                    // HalfC8: pixel px of the row -> staged halves 3 * (px - ix_first) ...
            const HalfC8* srcp = reinterpret_cast<const HalfC8*>(img) + ((size_t)b * H + iy) * W;
            const int p_lo = e_lo / 3, p_hi = (e_hi + 2) / 3;   // staged pixels [p_lo, p_hi)
            for (int pp = p_lo + tid; pp < p_hi; pp += STEM_THREADS) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(srcp + ix_first + pp));
                const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (pp * 3 + c < e_hi) dst[pp * 3 + c] = h[c];
            }
        } else {
        const TIn* src = base + (size_t)iy * W * 3;           // element e of the staged row = src[ix_first * 3 + e]
        const TIn* s0 = src + ix_first * 3 + e_lo;            // first valid source element
        const int n = e_hi - e_lo;
        constexpr int VEC = 4;                                // elements per vector: float4 (16 B) / uchar4 (4 B)
        const int mis = (int)(((uintptr_t)s0 / sizeof(TIn)) & (VEC - 1));
        const int head = min(n, (VEC - mis) & (VEC - 1));
        for (int e = tid; e < head; e += STEM_THREADS) dst[e_lo + e] = __float2half_rn(stem_to_float(s0[e]));
        const int nvec = (n - head) / VEC;
        if (sizeof(TIn) == 4) {
            const float4* v = reinterpret_cast<const float4*>(s0 + head);
            for (int i = tid; i < nvec; i += STEM_THREADS) {
                const float4 x = __ldg(v + i);
                __half* d = dst + e_lo + head + i * 4;
                d[0] = __float2half_rn(x.x); d[1] = __float2half_rn(x.y); d[2] = __float2half_rn(x.z); d[3] = __float2half_rn(x.w);
            }
        } else {
            const uchar4* v = reinterpret_cast<const uchar4*>(s0 + head);
            for (int i = tid; i < nvec; i += STEM_THREADS) {
                const uchar4 x = __ldg(v + i);
                __half* d = dst + e_lo + head + i * 4;
                d[0] = __float2half_rn(stem_to_float(x.x)); d[1] = __float2half_rn(stem_to_float(x.y));
                d[2] = __float2half_rn(stem_to_float(x.z)); d[3] = __float2half_rn(stem_to_float(x.w));
            }
        }
        for (int e = head + nvec * VEC + tid; e < n; e += STEM_THREADS) dst[e_lo + e] = __float2half_rn(stem_to_float(s0[e]));
        }
    }
    __syncthreads();

    // ---- 32 pixels x 32 channels per warp --------------------------------------------------------------------
    // A fragment element (row = pixel, k): k -> (ky = k / 9, j = k % 9) lives at srow[ky][STEP * pixel + j]
    int koff[8];                                              // staged offset of this thread's 8 k values (-1: zero)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = 16 * (q >> 2) + 8 * ((q >> 1) & 1) + 2 * t + (q & 1);
        koff[q] = k < 27 ? (k / 9) * STEM_ROWLEN + (k % 9) : -1;
    }
    const __half* sflat = &srow[0][0];
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[m][j][r] = 0.f;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int p0 = (warp * 32 + m * 16 + g) * STEP, p1 = p0 + 8 * STEP;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            uint32_t a[4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {                     // r: k half (k .. k+7 / k+8 .. k+15)
                const int q = s * 4 + r * 2;
                const uint16_t x00 = koff[q] >= 0 ? __half_as_ushort(sflat[koff[q] + p0]) : (uint16_t)0;
                const uint16_t x01 = koff[q + 1] >= 0 ? __half_as_ushort(sflat[koff[q + 1] + p0]) : (uint16_t)0;
                const uint16_t x10 = koff[q] >= 0 ? __half_as_ushort(sflat[koff[q] + p1]) : (uint16_t)0;
                const uint16_t x11 = koff[q + 1] >= 0 ? __half_as_ushort(sflat[koff[q + 1] + p1]) : (uint16_t)0;
                a[r * 2 + 0] = (uint32_t)x00 | ((uint32_t)x01 << 16);      // (row g,     k pair)
                a[r * 2 + 1] = (uint32_t)x10 | ((uint32_t)x11 << 16);      // (row g + 8, k pair)
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(acc[m][j], a, bf[s][j][0], bf[s][j][1]);
        }
    }

    // ---- bias + activation -> fp16 tile -> coalesced 16-byte stores ------------------------------------------
    const float lo = act == SSD_ACT_NONE ? -INFINITY : 0.0f, hi = act == SSD_ACT_RELU6 ? 6.0f : INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int c = 8 * j + 2 * t;
        const float b0 = bias ? __ldg(bias + c) : 0.f, b1 = bias ? __ldg(bias + c + 1) : 0.f;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int p = warp * 32 + m * 16 + g;
            const __half2 v0 = __floats2half2_rn(fminf(fmaxf(acc[m][j][0] + b0, lo), hi), fminf(fmaxf(acc[m][j][1] + b1, lo), hi));
            const __half2 v1 = __floats2half2_rn(fminf(fmaxf(acc[m][j][2] + b0, lo), hi), fminf(fmaxf(acc[m][j][3] + b1, lo), hi));
            *reinterpret_cast<__half2*>(&sout[p * STEM_OSTRIDE + c]) = v0;
            *reinterpret_cast<__half2*>(&sout[(p + 8) * STEM_OSTRIDE + c]) = v1;
        }
    }
    __syncthreads();
    const int npix = min(STEM_PIX, Wo - ox0);
    uint4* orow = reinterpret_cast<uint4*>(out + (((size_t)b * Ho + oy) * Wo + ox0) * COUT);
    for (int i = tid; i < npix * NT; i += STEM_THREADS)
        orow[i] = *reinterpret_cast<const uint4*>(&sout[(i / NT) * STEM_OSTRIDE + (i % NT) * 8]);
}

template <typename TIn>
static int stem_launch(const TIn* d_img, const void* d_weight, const float* d_bias, void* d_out, int B, int H, int W,
                       int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act, ssd_stream_t stream, const char* who) {
    SSD_REQUIRE_PTR(d_img); SSD_REQUIRE_PTR(d_weight); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Ho >= 1 && Wo >= 1 && act >= SSD_ACT_NONE && act <= SSD_ACT_RELU6 &&
                pad_top >= 0 && pad_left >= 0 && pad_top <= 1 && pad_left <= 1 && B <= 65535,
                SSD_ERR_SHAPE, "%s: bad shape B=%d H=%d W=%d Ho=%d Wo=%d act=%d pad=%d,%d", who, B, H, W, Ho, Wo, act, pad_top, pad_left);
    SSD_REQUIRE((stride == 2 && Cout == 32) || (stride == 1 && Cout == 64), SSD_ERR_UNSUPPORTED,
                "%s: stride=%d Cout=%d (this build instantiates stride 2 / Cout 32 and stride 1 / Cout 64)", who, stride, Cout);
    SSD_REQUIRE((Ho - 1) * stride - pad_top + 2 <= H && (Wo - 1) * stride - pad_left + 2 <= W,     // at most one padded row / column after
                SSD_ERR_SHAPE, "%s: output %dx%d does not fit input %dx%d with stride %d", who, Ho, Wo, H, W, stride);
    const int chunks = ceil_div(Wo, STEM_PIX);
    auto kern = stride == 2 ? stem_conv3x3s2_mma_kernel<TIn, 2, 4> : stem_conv3x3s2_mma_kernel<TIn, 1, 8>;
    cudaError_t le = launch_pdl(kern, dim3(Ho * chunks, B), dim3(STEM_THREADS), 0, as_stream(stream),
                                d_img, reinterpret_cast<const __half*>(d_weight), d_bias, reinterpret_cast<__half*>(d_out),
                                H, W, Ho, Wo, pad_top, pad_left, act, chunks);
    if (le != cudaSuccess) return cuda_fail(le, "stem_conv3x3s2_mma_kernel");
    return SSD_OK;
}

// uint8 NHWC (3 channels) -> fp16 NHWC with the channel axis padded to 8: convert_image_dtype fused into the cast
// that feeds a tensor-core first layer (VGG16 conv1_1).
__global__ void __launch_bounds__(256)
image_u8_to_f16c8_kernel(const uint8_t* __restrict__ img, uint4* __restrict__ out, int64_t n_pixels) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        __half2* h = reinterpret_cast<__half2*>(&o);
        h[0] = __floats2half2_rn(stem_to_float(img[i * 3]), stem_to_float(img[i * 3 + 1]));
        h[1] = __floats2half2_rn(stem_to_float(img[i * 3 + 2]), 0.f);
        out[i] = o;
    }
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_stem_conv3x3s2(const float* d_img, const void* d_weight, const float* d_bias, void* d_out,
                                  int B, int H, int W, int Cout, int Ho, int Wo, int pad_top, int pad_left, int act,
                                  ssd_stream_t stream) {
    return stem_launch<float>(d_img, d_weight, d_bias, d_out, B, H, W, Cout, Ho, Wo, 2, pad_top, pad_left, act, stream,
                              "ssd_stem_conv3x3s2");
}

extern "C" int ssd_stem_conv3x3(const float* d_img, const void* d_weight, const float* d_bias, void* d_out,
                                int B, int H, int W, int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act,
                                ssd_stream_t stream) {
    return stem_launch<float>(d_img, d_weight, d_bias, d_out, B, H, W, Cout, Ho, Wo, stride, pad_top, pad_left, act, stream,
                              "ssd_stem_conv3x3");
}

extern "C" int ssd_stem_conv3x3_u8(const void* d_img_u8, const void* d_weight, const float* d_bias, void* d_out,
                                   int B, int H, int W, int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act,
                                   ssd_stream_t stream) {
    return stem_launch<uint8_t>(static_cast<const uint8_t*>(d_img_u8), d_weight, d_bias, d_out, B, H, W, Cout, Ho, Wo,
                                stride, pad_top, pad_left, act, stream, "ssd_stem_conv3x3_u8");
}

extern "C" int ssd_stem_conv3x3s2_u8(const void* d_img_u8, const void* d_weight, const float* d_bias, void* d_out,
                                     int B, int H, int W, int Cout, int Ho, int Wo, int pad_top, int pad_left, int act,
                                     ssd_stream_t stream) {
    return stem_launch<uint8_t>(static_cast<const uint8_t*>(d_img_u8), d_weight, d_bias, d_out, B, H, W, Cout, Ho, Wo,
                                2, pad_top, pad_left, act, stream, "ssd_stem_conv3x3s2_u8");
}

extern "C" int ssd_stem_conv3x3_f16c8(const void* d_img_f16c8, const void* d_weight_ohwi8, const float* d_bias, void* d_out,
                                      int B, int H, int W, int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act,
                                      ssd_stream_t stream) {
    return stem_launch<HalfC8>(static_cast<const HalfC8*>(d_img_f16c8), d_weight_ohwi8, d_bias, d_out, B, H, W, Cout, Ho, Wo,
                               stride, pad_top, pad_left, act, stream, "ssd_stem_conv3x3_f16c8");
}

extern "C" int ssd_image_u8_to_f16c8(const void* d_img_u8, void* d_out, int64_t n_pixels, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_img_u8); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(n_pixels >= 0, SSD_ERR_SHAPE, "ssd_image_u8_to_f16c8: n_pixels=%lld", (long long)n_pixels);
    if (n_pixels == 0) return SSD_OK;
    const int64_t blocks = (n_pixels + 255) / 256, cap = (int64_t)sm_count() * 16;
    image_u8_to_f16c8_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        static_cast<const uint8_t*>(d_img_u8), reinterpret_cast<uint4*>(d_out), n_pixels);
    SSD_CHECK_LAUNCH("image_u8_to_f16c8_kernel");
    return SSD_OK;
}
