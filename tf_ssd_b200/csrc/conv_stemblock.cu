// MobileNetV2's first three layers as ONE launch on sm_100a:
//
//     Conv1 3x3 stride 2 (3 -> 32, + folded bn_Conv1 + ReLU6)            keras_applications Conv1_pad / Conv1 / bn_Conv1
//  -> expanded_conv_depthwise 3x3 stride 1 (+ folded BN + ReLU6)         (block 0 has no 1x1 expansion)
//  -> expanded_conv_project 1x1 (32 -> 16, + folded BN)                  (models/ssd_mobilenet_v2.py:25 of the reference)
//
// As separate launches the 150x150x32 stem output (46 MB per batch of 32) is written, evicted from L2 and read back,
// and block 0 runs the tcgen05 depthwise->projection pipeline at one 128-pixel tile per ~2 us (86 us + 40 us for the
// stem).  None of the three layers has a deep contraction (K = 27, 9, 32), so this kernel keeps them on mma.sync / packed
// half2 FMAs inside ONE CTA per 30 x 10 output tile, with plain __syncthreads() between phases and four CTAs (54 KB, 64 registers)
// per SM overlapping each other's phases:
//
//   1. the 25 x 65 image patch (u8 or f32) is staged in shared memory as fp16 (convert_image_dtype fused in);
//   2. stem: the 32 x 12 halo patch of Conv1's output = 24 m16 tiles x (K = 27 -> 32) x 32 channels on mma.sync, A
//      fragments gathered from the staged rows; + bias, ReLU6, fp16, ZERO outside the map (the depthwise padding), into a
//      swizzled [position][32 ch] tile;
//   3. depthwise 3x3: a thread owns 8 channels x 5 horizontally adjacent pixels (sliding window: 21 vector loads per 5
//      outputs, the 9 x 8 filter taps live in registers), packed half2 FMAs like ssd_dwproj / ssd_irblock;
//   4. projection: 19 m16 tiles x (K = 32) x Cout on mma.sync from the depthwise tile; + bias -> fp16 staging tile;
//   5. 16-byte coalesced stores of the 30 x 10 x Cout tile.
//
// HBM traffic: the image once (+ halo re-reads through L2) and the 150x150x16 output once: 32 MB instead of 124 MB.

#include "common.cuh"

#include <type_traits>

namespace ssd {

constexpr int SB_TW = 30, SB_TH = 10;                 // output tile: 150 = 5 x 30 = 15 x 10 (no partial tiles at SSD300)
constexpr int SB_PW = SB_TW + 2, SB_PH = SB_TH + 2;   // halo patch of the stem output: 32 x 12
constexpr int SB_NPOS = SB_PW * SB_PH;                // 384 positions = 24 m16 tiles
constexpr int SB_NPIX = SB_TW * SB_TH;                // 300 pixels -> 19 m16 tiles (304 rows)
constexpr int SB_MT_STEM = SB_NPOS / 16;
constexpr int SB_MT_PROJ = (SB_NPIX + 15) / 16;
constexpr int SB_IROWS = 2 * SB_PH + 1;               // 25 image rows
constexpr int SB_IROWLEN = 200;                       // staged halves per row: 65 x 3 = 195, + up to 3 of misalignment
constexpr int SB_THREADS = 256;
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_DW_XG = SB_TW / 5;                   // groups of 5 adjacent pixels per tile row
constexpr int SB_DW_ITEMS = SB_TH * SB_DW_XG * 4;     // (row, group, 8-channel chunk) = 240 threads
static_assert(SB_PW == 32, "positions decompose with shifts");
static_assert(SB_TW % 5 == 0 && SB_DW_ITEMS <= SB_THREADS, "depthwise mapping");
constexpr int SB_OFF_IMG = 0;
constexpr int SB_OFF_MID = SB_IROWS * SB_IROWLEN * 2;                   // 10 000
constexpr int SB_OFF_DW = SB_OFF_MID + SB_NPOS * 64;                    // + 24 576
constexpr int SB_SMEM = SB_OFF_DW + SB_MT_PROJ * 16 * 64;               // + 19 456 = 54 032
static_assert(SB_OFF_MID % 16 == 0 && SB_OFF_DW % 16 == 0, "16-byte aligned tiles");
static_assert(SB_NPIX * 64 <= SB_NPOS * 64, "the output staging tile aliases the stem tile");

__device__ __forceinline__ float sb_to_float(float v) { return v; }
__device__ __forceinline__ float sb_to_float(uint8_t v) { return __fmul_rn((float)v, 1.0f / 255.0f); }

__device__ __forceinline__ void sb_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// [row][32 fp16 channels] tiles, 64 bytes per row, the 16-byte chunk index XOR-ed with (row >> 1) & 3: the 4-byte
// accumulator-fragment stores (8 consecutive rows x 4 lanes), the 16-byte depthwise loads (4 chunks x 2 rows of
// different parity per quarter warp) and the 4-byte A-fragment loads of the projection are all bank-conflict free.
__device__ __forceinline__ uint32_t sb_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint4 sb_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t sb_lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sb_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sb_sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t sb_pack(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

struct SbParams {
    const void* img; const __half* w_stem; const float* b_stem; const __half* w_dw; const float* b_dw;
    const __half* w_proj; const float* b_proj; __half* out;
    int H, W, Hs, Ws, pad_t, pad_l, tiles_x, stem_act, dw_act, act;
};

template <typename TIn, int NTP>
__global__ void __launch_bounds__(SB_THREADS, 4)
stem_dwproj_kernel(const __grid_constant__ SbParams p) {
    extern __shared__ __align__(16) unsigned char sb_smem[];
    __half* sImg = reinterpret_cast<__half*>(sb_smem + SB_OFF_IMG);
    const uint32_t sImg32 = (uint32_t)__cvta_generic_to_shared(sb_smem + SB_OFF_IMG);
    const uint32_t sMid = (uint32_t)__cvta_generic_to_shared(sb_smem + SB_OFF_MID);
    const uint32_t sDw = (uint32_t)__cvta_generic_to_shared(sb_smem + SB_OFF_DW);
    constexpr int COUT = 8 * NTP;
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x, b = blockIdx.y;
    const int sy0 = ty * SB_TH - 1, sx0 = tx * SB_TW - 1;         // stem-output coordinates of patch position (0, 0)
    const int iy0 = 2 * sy0 - p.pad_t;                            // image row of staged row 0
    const int ebase = (2 * sx0 - p.pad_l) * 3;                    // image-row element (x * 3 + c) of patch column 0, channel 0
    const int shift = ebase & 3;                                  // staged rows start at the 4-element boundary below it
    const int abase = ebase - shift;
    const int rowlen = p.W * 3;

    // The stem's K axis (27 taps, padded to 32) is ordered in 16 PAIR SLOTS so that every A-fragment register is ONE
    // aligned 32-bit shared-memory load: slot ky * 5 + i holds the staged elements (jstart, jstart + 1) of image row
    // 2 pr + ky with jstart = 2 i - (shift & 1) (j = kx * 3 + ci in 0..8; the elements outside that range belong to
    // the neighbouring pixel and meet a zero weight); slot 15 is all zero.  Thread t of a quad owns slots t, 4 + t,
    // 8 + t, 12 + t (mma.sync's k = 2t, 2t+1 / 2t+8, 2t+9 of the two k16 steps).
    const int sodd = shift & 1;
    uint32_t bf[2][4][2];
    int soff[4];                                                  // staged offset (halves) of this thread's four slots
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int slot = 8 * s + 4 * r + t;
            const int sl = slot < 15 ? slot : 14;
            const int ky = sl / 5, jstart = 2 * (sl - ky * 5) - sodd;
            soff[2 * s + r] = ky * SB_IROWLEN + shift + jstart;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half* wr = p.w_stem + (8 * j + g) * 27 + ky * 9;
                const __half lo = (slot < 15 && jstart >= 0) ? wr[jstart] : __float2half(0.f);
                const __half hi = (slot < 15 && jstart + 1 <= 8) ? wr[jstart + 1] : __float2half(0.f);
                bf[s][j][r] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
        }
    pdl_wait();

    // ---- 1. the image patch as fp16: staged element s of row r = image element abase + s of image row iy0 + r ----
    {
        const TIn* img = static_cast<const TIn*>(p.img) + (size_t)b * p.H * rowlen;
        constexpr int VPR = SB_IROWLEN / 4, NVEC = SB_IROWS * VPR;      // 50 four-element vectors per row, 1250 per patch
        if ((rowlen & 3) == 0) {
            // rows and abase are multiples of four elements: a vector is inside the image or outside, never across
            // its edge.  All loads of a thread are issued before the first conversion.
            constexpr int IT = (NVEC + SB_THREADS - 1) / SB_THREADS;
            static_assert(SB_THREADS / VPR == 5 && SB_THREADS % VPR == 6, "incremental (row, vector) update below");
            typedef typename std::conditional<sizeof(TIn) == 4, float4, uchar4>::type VecT;
            VecT q[IT];
            int r = tid / VPR, v = tid - r * VPR;
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                const int iy = iy0 + r, e0 = abase + 4 * v;
                const bool ok = tid + it * SB_THREADS < NVEC && (unsigned)iy < (unsigned)p.H && (unsigned)e0 < (unsigned)rowlen;
                if constexpr (sizeof(TIn) == 4) q[it] = make_float4(0.f, 0.f, 0.f, 0.f); else q[it] = make_uchar4(0, 0, 0, 0);
                if (ok) q[it] = __ldg(reinterpret_cast<const VecT*>(img + (size_t)iy * rowlen + e0));
                v += SB_THREADS % VPR; r += SB_THREADS / VPR;
                if (v >= VPR) { v -= VPR; ++r; }
            }
            r = tid / VPR; v = tid - r * VPR;
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                uint2 o;
                o.x = sb_pack(__floats2half2_rn(sb_to_float(q[it].x), sb_to_float(q[it].y)));
                o.y = sb_pack(__floats2half2_rn(sb_to_float(q[it].z), sb_to_float(q[it].w)));
                if (tid + it * SB_THREADS < NVEC) *reinterpret_cast<uint2*>(sImg + r * SB_IROWLEN + 4 * v) = o;
                v += SB_THREADS % VPR; r += SB_THREADS / VPR;
                if (v >= VPR) { v -= VPR; ++r; }
            }
        } else {
            for (int i = tid; i < NVEC; i += SB_THREADS) {
                const int r = i / VPR, v = i - r * VPR;
                const int iy = iy0 + r, e0 = abase + 4 * v;
                float x[4] = {0.f, 0.f, 0.f, 0.f};
                if ((unsigned)iy < (unsigned)p.H) {
                    const TIn* src = img + (size_t)iy * rowlen;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (e0 + c >= 0 && e0 + c < rowlen) x[c] = sb_to_float(__ldg(src + e0 + c));
                }
                uint2 o;
                o.x = sb_pack(__floats2half2_rn(x[0], x[1]));
                o.y = sb_pack(__floats2half2_rn(x[2], x[3]));
                *reinterpret_cast<uint2*>(sImg + r * SB_IROWLEN + 4 * v) = o;
            }
        }
    }
    __syncthreads();

    // ---- 2. stem: patch positions x 32 channels; position = pr * 32 + pc, tap (ky, j = kx * 3 + ci) of a position lives at
    //         sImg[(2 pr + ky) * ROWLEN + 6 pc + shift + j] ---------------------------------------------------------
    {
        const __half2 lo2 = __float2half2_rn(p.stem_act == SSD_ACT_NONE ? -65504.0f : 0.0f);
        const __half2 hi2 = __float2half2_rn(p.stem_act == SSD_ACT_RELU6 ? 6.0f : 65504.0f);
        float bias[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bias[j][0] = p.b_stem ? __ldg(p.b_stem + 8 * j + 2 * t) : 0.f;
            bias[j][1] = p.b_stem ? __ldg(p.b_stem + 8 * j + 2 * t + 1) : 0.f;
        }
        for (int mt = warp; mt < SB_MT_STEM; mt += SB_WARPS) {
            const int pos0 = mt * 16 + g;                          // rows g and g + 8 of the m16 tile: same patch row
            const int pr = pos0 >> 5, pc = pos0 & 31;
            const int p0 = 2 * pr * SB_IROWLEN + 6 * pc, p1 = p0 + 48;
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[j][r] = 0.f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t a[4];
                a[0] = sb_lds32(sImg32 + (uint32_t)(p0 + soff[2 * s]) * 2u);
                a[1] = sb_lds32(sImg32 + (uint32_t)(p1 + soff[2 * s]) * 2u);
                a[2] = sb_lds32(sImg32 + (uint32_t)(p0 + soff[2 * s + 1]) * 2u);
                a[3] = sb_lds32(sImg32 + (uint32_t)(p1 + soff[2 * s + 1]) * 2u);
#pragma unroll
                for (int j = 0; j < 4; ++j) sb_mma16816(acc[j], a, bf[s][j][0], bf[s][j][1]);
            }
            // positions outside the stem output are the depthwise layer's zero padding
            const int sy = sy0 + pr, sx = sx0 + pc;
            const bool row_ok = (unsigned)sy < (unsigned)p.Hs;
            const uint32_t keep0 = (row_ok && (unsigned)sx < (unsigned)p.Ws) ? 0xffffffffu : 0u;
            const uint32_t keep1 = (row_ok && (unsigned)(sx + 8) < (unsigned)p.Ws) ? 0xffffffffu : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 v0 = __hmin2(__hmax2(__floats2half2_rn(acc[j][0] + bias[j][0], acc[j][1] + bias[j][1]), lo2), hi2);
                const __half2 v1 = __hmin2(__hmax2(__floats2half2_rn(acc[j][2] + bias[j][0], acc[j][3] + bias[j][1]), lo2), hi2);
                sb_sts32(sMid + sb_off(pos0, j) + 4 * t, sb_pack(v0) & keep0);
                sb_sts32(sMid + sb_off(pos0 + 8, j) + 4 * t, sb_pack(v1) & keep1);
            }
        }
    }
    __syncthreads();

    // ---- 3. depthwise 3x3 (stride 1, zero padding already in the tile): 8 channels x 5 adjacent pixels per thread ----
    if (tid < SB_DW_ITEMS) {
        const int j = tid & 3, xg = (tid >> 2) % SB_DW_XG, y = tid / (4 * SB_DW_XG);
        uint4 w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = __ldg(reinterpret_cast<const uint4*>(p.w_dw + k * 32 + 8 * j));
        __half2 bias4[4];
        {
            const float4 b0 = p.b_dw ? __ldg(reinterpret_cast<const float4*>(p.b_dw + 8 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = p.b_dw ? __ldg(reinterpret_cast<const float4*>(p.b_dw + 8 * j + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            bias4[0] = __floats2half2_rn(b0.x, b0.y); bias4[1] = __floats2half2_rn(b0.z, b0.w);
            bias4[2] = __floats2half2_rn(b1.x, b1.y); bias4[3] = __floats2half2_rn(b1.z, b1.w);
        }
        // packed half2 FMAs: the nine products of an output are summed in fp16 starting from the fp16-rounded bias
        // (the arithmetic of ssd_dwproj / ssd_irblock)
        __half2 acc[5][4];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = bias4[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const uint4 xv = sb_lds128(sMid + sb_off((y + ky) * SB_PW + 5 * xg + c, j));
                const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int kx = c - i;
                    if (kx >= 0 && kx < 3) {
                        const __half2* wh = reinterpret_cast<const __half2*>(&w[ky * 3 + kx]);
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) acc[i][c2] = __hfma2(xh[c2], wh[c2], acc[i][c2]);
                    }
                }
            }
        const __half2 lo2 = __float2half2_rn(p.dw_act == SSD_ACT_NONE ? -65504.0f : 0.0f);
        const __half2 hi2 = __float2half2_rn(p.dw_act == SSD_ACT_RELU6 ? 6.0f : 65504.0f);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) oh[c2] = __hmin2(__hmax2(acc[i][c2], lo2), hi2);
            sb_sts128(sDw + sb_off(y * SB_TW + 5 * xg + i, j), o);
        }
    }
    __syncthreads();

    // ---- 4. projection 1x1: [pixels x 32] x [32 x Cout] on mma.sync; the staging tile aliases the stem tile ----
    {
        uint32_t pf[2][NTP][2];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int n = 0; n < NTP; ++n)
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    pf[s][n][r] = __ldg(reinterpret_cast<const uint32_t*>(p.w_proj + (8 * n + g) * 32 + 16 * s + 8 * r + 2 * t));
        float pb[NTP][2];
#pragma unroll
        for (int n = 0; n < NTP; ++n) {
            pb[n][0] = p.b_proj ? __ldg(p.b_proj + 8 * n + 2 * t) : 0.f;
            pb[n][1] = p.b_proj ? __ldg(p.b_proj + 8 * n + 2 * t + 1) : 0.f;
        }
        const float lo = p.act == SSD_ACT_NONE ? -INFINITY : 0.0f, hi = p.act == SSD_ACT_RELU6 ? 6.0f : INFINITY;
        const uint32_t sOut = sMid;
        for (int mt = warp; mt < SB_MT_PROJ; mt += SB_WARPS) {
            const int px0 = mt * 16 + g;
            float acc[NTP][4];
#pragma unroll
            for (int n = 0; n < NTP; ++n)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[n][r] = 0.f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t a[4];
                a[0] = sb_lds32(sDw + sb_off(px0, 2 * s) + 4 * t);
                a[1] = sb_lds32(sDw + sb_off(px0 + 8, 2 * s) + 4 * t);
                a[2] = sb_lds32(sDw + sb_off(px0, 2 * s + 1) + 4 * t);
                a[3] = sb_lds32(sDw + sb_off(px0 + 8, 2 * s + 1) + 4 * t);
#pragma unroll
                for (int n = 0; n < NTP; ++n) sb_mma16816(acc[n], a, pf[s][n][0], pf[s][n][1]);
            }
#pragma unroll
            for (int n = 0; n < NTP; ++n) {
                const __half2 v0 = __floats2half2_rn(fminf(fmaxf(acc[n][0] + pb[n][0], lo), hi), fminf(fmaxf(acc[n][1] + pb[n][1], lo), hi));
                const __half2 v1 = __floats2half2_rn(fminf(fmaxf(acc[n][2] + pb[n][0], lo), hi), fminf(fmaxf(acc[n][3] + pb[n][1], lo), hi));
                if (px0 < SB_NPIX) sb_sts32(sOut + (uint32_t)(px0 * COUT + 8 * n + 2 * t) * 2u, sb_pack(v0));
                if (px0 + 8 < SB_NPIX) sb_sts32(sOut + (uint32_t)((px0 + 8) * COUT + 8 * n + 2 * t) * 2u, sb_pack(v1));
            }
        }
    }
    __syncthreads();

    // ---- 5. the tile's rows are contiguous runs of the NHWC output: 16-byte coalesced stores ----
    {
        const int nx = min(SB_TW, p.Ws - tx * SB_TW);
        uint4* out16 = reinterpret_cast<uint4*>(p.out);
        for (int i = tid; i < SB_NPIX * NTP; i += SB_THREADS) {
            const int y = i / (SB_TW * NTP), rem = i - y * (SB_TW * NTP), x = rem / NTP, u = rem - x * NTP;
            const int sy = ty * SB_TH + y;
            if (sy < p.Hs && x < nx)
                out16[((size_t)(b * p.Hs + sy) * p.Ws + tx * SB_TW + x) * NTP + u] = sb_lds128(sMid + (uint32_t)i * 16u);
        }
    }
}

static bool stem_dwproj_ok(const ssd_stem_dwproj_desc* d) {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    return d->Cmid == 32 && d->Cout >= 8 && d->Cout <= 32 && d->Cout % 8 == 0 && d->B >= 1 && d->B <= 65535 &&
           d->H >= 1 && d->W >= 1 && d->Hs >= 1 && d->Ws >= 1 && d->pad_top >= 0 && d->pad_top <= 1 && d->pad_left >= 0 &&
           d->pad_left <= 1 && (d->Hs - 1) * 2 - d->pad_top + 2 <= d->H && (d->Ws - 1) * 2 - d->pad_left + 2 <= d->W &&
           (d->image_u8 == 0 || d->image_u8 == 1) && al16(d->image) && al16(d->dw_weight) && al16(d->proj_weight) &&
           al16(d->out) && (d->dw_bias == nullptr || al16(d->dw_bias)) &&
           d->stem_act >= SSD_ACT_NONE && d->stem_act <= SSD_ACT_RELU6 && d->dw_act >= SSD_ACT_NONE && d->dw_act <= SSD_ACT_RELU6 &&
           d->act >= SSD_ACT_NONE && d->act <= SSD_ACT_RELU6;
}

template <typename TIn, int NTP>
static cudaError_t stem_dwproj_launch_t(const SbParams& p, dim3 grid, cudaStream_t st) {
    static thread_local int attr_dev = -1;
    int cur = 0;
    cudaGetDevice(&cur);
    if (attr_dev != cur) {
        cudaError_t e = cudaFuncSetAttribute(stem_dwproj_kernel<TIn, NTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM);
        if (e != cudaSuccess) return e;
        attr_dev = cur;
    }
    return launch_pdl(stem_dwproj_kernel<TIn, NTP>, grid, dim3(SB_THREADS), (size_t)SB_SMEM, st, p);
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_stem_dwproj_supported(const ssd_stem_dwproj_desc* d) {
    if (!d || !d->image || !d->stem_weight || !d->dw_weight || !d->proj_weight || !d->out) return 0;
    return stem_dwproj_ok(d) ? 1 : 0;
}

extern "C" int ssd_stem_dwproj(const ssd_stem_dwproj_desc* d, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d);
    SSD_REQUIRE_PTR(d->image); SSD_REQUIRE_PTR(d->stem_weight); SSD_REQUIRE_PTR(d->dw_weight); SSD_REQUIRE_PTR(d->proj_weight);
    SSD_REQUIRE_PTR(d->out);
    SSD_REQUIRE(stem_dwproj_ok(d), SSD_ERR_UNSUPPORTED,
                "ssd_stem_dwproj: unsupported configuration (Cmid == 32, Cout in 8..32 step 8, stride-2 stem with pads <= 1, "
                "16-byte aligned pointers): B=%d H=%d W=%d Hs=%d Ws=%d Cmid=%d Cout=%d", d->B, d->H, d->W, d->Hs, d->Ws, d->Cmid, d->Cout);
    SbParams p;
    p.img = d->image; p.w_stem = static_cast<const __half*>(d->stem_weight); p.b_stem = d->stem_bias;
    p.w_dw = static_cast<const __half*>(d->dw_weight); p.b_dw = d->dw_bias;
    p.w_proj = static_cast<const __half*>(d->proj_weight); p.b_proj = d->proj_bias; p.out = static_cast<__half*>(d->out);
    p.H = d->H; p.W = d->W; p.Hs = d->Hs; p.Ws = d->Ws; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
    p.tiles_x = ceil_div(d->Ws, SB_TW);
    p.stem_act = d->stem_act; p.dw_act = d->dw_act; p.act = d->act;
    const dim3 grid(p.tiles_x * ceil_div(d->Hs, SB_TH), d->B);
    cudaStream_t st = as_stream(stream);
    cudaError_t e;
    const int ntp = d->Cout / 8;
#define SB_CASE(N) (d->image_u8 ? stem_dwproj_launch_t<uint8_t, N>(p, grid, st) : stem_dwproj_launch_t<float, N>(p, grid, st))
    e = ntp == 1 ? SB_CASE(1) : ntp == 2 ? SB_CASE(2) : ntp == 3 ? SB_CASE(3) : SB_CASE(4);
#undef SB_CASE
    if (e != cudaSuccess) return cuda_fail(e, "stem_dwproj_kernel");
    return SSD_OK;
}
