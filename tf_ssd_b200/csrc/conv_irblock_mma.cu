// The LARGE-MAP inverted-residual blocks of MobileNetV2 (blocks 1-6: 150x150 ... 38x38, Cin <= 32, Cexp <= 192) as one
// launch each, on mma.sync -- second implementation behind ssd_irblock (keras_applications.mobilenet_v2.
// _inverted_res_block under models/ssd_mobilenet_v2.py:25 of the reference):
//
//     1x1 expand (+ folded BN + ReLU6) -> depthwise 3x3 stride 1|2 (+ folded BN + ReLU6) -> 1x1 project (+ folded BN, + shortcut)
//
// Why not the tcgen05 pipeline of conv_irblock.cu here: these blocks have shallow contractions (K = 16..32 for the
// expansion, N = 24..64 for the projection) and small weights, so the tensor pipe is idle either way; the tcgen05 kernel
// is ONE CTA per SM whose five roles hand 64-channel slices over through mbarriers, and its slice period (~2 us) is a
// latency chain (TMA -> MMA -> TMEM load -> shared memory -> depthwise -> MMA), not work.  Here a CTA owns one small
// output tile with ALL expanded channels, phases are separated by plain __syncthreads(), and two CTAs per SM overlap each
// other's phases; the cost is issue slots (the expansion's epilogue and the depthwise FMAs), nothing waits on a chain.
//
//   1. TMA: expansion / projection weights once per (persistent) CTA, the input patch (tile + halo, zero outside the image)
//      per tile -- the NEXT tile's patch is requested as soon as the expansion has read the buffer; depthwise filter and
//      biases by the threads, once per CTA;
//   2. expansion: [patch positions x Cin] x [Cin x Cexp] on mma.sync.m16n8k16 (ldmatrix operands; the bias enters as the C
//      operand of the first MMA), ReLU6 fused into the fp16 conversion (cvt.rn.relu + one min) -> expanded patch in shared
//      memory; edge tiles then zero the positions outside the image (the depthwise layer's zero padding);
//   3. depthwise 3x3: a thread owns one 8-channel chunk (its 9 x 8 taps live in registers) and walks over tile pixels
//      (runs of three adjacent pixels at stride 1), packed half2 FMAs -> the projection's A operand in shared memory;
//   4. projection: [pixels x Cexp] x [Cexp x Cout] on mma.sync, + bias -> fp32 staging tile;
//   5. (+ residual) -> fp16 -> 16-byte coalesced stores.
//
// Shared-memory tiles are arrays of 16-channel SLICES: tile[slice][row][16 fp16] (32-byte rows); a k16 step of either GEMM
// is one slice.  The ldmatrix operands (input patch, weights, depthwise tile) XOR the 16-byte half of a row with
// (row >> 2) & 1 -- conflict-free 8-row phases, and exactly the tensor map's 32-byte swizzle, so TMA writes them directly.
// The expanded patch is linear (the nine depthwise taps of a pixel are compile-time offsets from one address) with one
// padding row per slice (the depthwise stage's 16-byte accesses to 8 consecutive chunks of a position = 4 slices are then
// bank-conflict free); inside a slice its 16 channels are stored in the order of the accumulator fragments (a thread's four
// values are contiguous: one 8-byte store per row), and the depthwise filter / biases are kept in that order too.
//
// The second kernel of this file (irblock_mma_grouped_kernel) runs the SMALL-MAP blocks 7-12, 14, 15 with the expanded
// channels in groups of 96 and the weights streamed through a TMA double buffer.

#include "tc_common.cuh"

#include <stdlib.h>

namespace ssd {

constexpr int IM_THREADS = 256;
constexpr int IM_WARPS = IM_THREADS / 32;

__host__ __device__ constexpr int im_up16(int v) { return (v + 15) / 16 * 16; }
__host__ __device__ constexpr int im_max(int a, int b) { return a > b ? a : b; }

template <int CIN_, int CEXP_, int COUT_, int STRIDE_, int TW_, int TH_>
struct ImCfg {
    static constexpr int CIN = CIN_, CEXP = CEXP_, COUT = COUT_, S = STRIDE_, TW = TW_, TH = TH_;
    static constexpr int KIN = im_up16(CIN);                  // expansion K, zero-padded to whole k16 steps
    static constexpr int KS = KIN / 16;                       // input slices = k16 steps of the expansion
    static constexpr int ES = CEXP / 16;                      // expanded slices = n-tile pairs of the expansion = k16 steps of the projection
    static constexpr int PW = (TW - 1) * S + 3, PH = (TH - 1) * S + 3;
    static constexpr int P = PW * PH, PPOS = im_up16(P);      // patch positions (rows of the expansion GEMM)
    static constexpr int NPX = TW * TH, MPX = im_up16(NPX);   // tile pixels (rows of the projection GEMM)
    static constexpr int COUTP = im_up16(COUT);               // projection N, whole n-tile pairs
    static constexpr int MT = PPOS / 16, MTP = MPX / 16, NP = COUTP / 16;
    static constexpr int NCH = CEXP / 8;                      // 8-channel chunks of the expanded tensor
    static constexpr int NPL = IM_THREADS / NCH;              // pixel lanes of the depthwise stage
    // slice strides (bytes).  TMA destinations (input patch, expansion / projection weights): whole rows, 128-byte aligned,
    // so that the tensor map's 32-byte swizzle == im_row()'s XOR.  Expanded patch and depthwise tile: one padding row per
    // slice (see the header comment).
    static constexpr int IN_SL = PPOS * 32, WE_SL = CEXP * 32, WP_SL = COUTP * 32;
    static constexpr int MID_SL = (PPOS + 1) * 32, DW_SL = (MPX + 1) * 32;
    // layout (256-byte aligned TMA destinations so that the tensor map's 32-byte swizzle, which XORs address bit 7 into bit 4,
    // equals im_row()'s XOR with (row >> 2) & 1).  The kernel is persistent: the weights are loaded once per CTA, and the
    // input patch of the NEXT tile is requested as soon as the expansion of the current one has read the buffer, so it lands
    // under the depthwise / projection / store phases.  The expanded patch is reused as the fp32 output staging tile.
    static constexpr int OFF_IN = 0;
    static constexpr int OFF_WE = OFF_IN + KS * IN_SL;
    static constexpr int OFF_WP = OFF_WE + KS * WE_SL;
    static constexpr int OFF_DW = OFF_WP + ES * WP_SL;
    static constexpr int OFF_MID = OFF_DW + (ES * DW_SL + 255) / 256 * 256;
    static constexpr int OFF_WD = OFF_MID + (im_max(ES * MID_SL, NPX * COUT * 4) + 255) / 256 * 256;   // [9][CEXP] fp16
    static constexpr int OFF_BE = OFF_WD + 9 * CEXP * 2;      // expansion bias [CEXP] f32, in the expanded patch's channel order
    static constexpr int OFF_BD = OFF_BE + CEXP * 4;          // depthwise bias [CEXP] f32
    static constexpr int OFF_BP = OFF_BD + CEXP * 4;          // projection bias [COUTP] f32
    static constexpr int OFF_BAR = OFF_BP + COUTP * 4;        // two mbarriers: weights (once), input patch (one phase per tile)
    static constexpr int SMEM = OFF_BAR + 16 + 1024;          // + slack for the manual 1024-byte alignment of the base
    static constexpr uint32_t W_BYTES = KS * WE_SL + ES * WP_SL, X_BYTES = KS * P * 32;
    static_assert(OFF_WE % 256 == 0 && OFF_WP % 256 == 0 && IN_SL % 256 == 0 && WE_SL % 256 == 0 && WP_SL % 256 == 0, "TMA destinations");
    static_assert(PW <= 256 && PH <= 256 && CEXP <= 256 && COUTP <= 256, "TMA box dimensions");
    static_assert(CIN % 8 == 0 && CEXP % 16 == 0 && COUT % 8 == 0, "channel granularity");
    static_assert(OFF_DW % 16 == 0 && OFF_MID % 16 == 0 && OFF_WD % 16 == 0 && OFF_BE % 16 == 0 && OFF_BD % 16 == 0 && OFF_BP % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
    static_assert(2 * (SMEM + 1024) <= 227 * 1024, "two CTAs per SM");
    static_assert(NCH <= IM_THREADS, "depthwise mapping");
};

struct ImParams {
    const __half* in; const __half* we; const float* be; const __half* wd; const float* bd;
    const __half* wp; const float* bp; const __half* res; __half* out;
    int B, H, W, Ho, Wo, pad_t, pad_l, tiles_x, tiles_per_img, n_tiles;
};

// byte offset of (row, 16-byte half h) inside a slice
__device__ __forceinline__ uint32_t im_row(int row, int h) { return (uint32_t)(row * 32 + ((h ^ ((row >> 2) & 1)) << 4)); }
__device__ __forceinline__ void im_ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void im_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// first k-step of the expansion: the accumulator starts at the bias (C = {c0, c1, c0, c1}: rows g and g + 8 share the columns)
__device__ __forceinline__ void im_mma_bias(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
}
// fp16x2(max(lo, 0), max(hi, 0)): the lower bound of ReLU6 fused into the conversion
__device__ __forceinline__ uint32_t im_cvt_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t im_min2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint4 im_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void im_sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void im_sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void im_sts64f(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 im_lds64f(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 im_lds128f(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

template <class Cfg>
__global__ void __launch_bounds__(IM_THREADS, 2)
irblock_mma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_we,
                   const __grid_constant__ CUtensorMap map_wp, const __grid_constant__ ImParams p) {
    extern __shared__ __align__(16) unsigned char im_smem_raw[];
    unsigned char* im_smem = reinterpret_cast<unsigned char*>(((uintptr_t)im_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(im_smem);
    const uint32_t sIn = sm + Cfg::OFF_IN, sWe = sm + Cfg::OFF_WE, sDw = sm + Cfg::OFF_DW, sMid = sm + Cfg::OFF_MID;
    const uint32_t sWp = sm + Cfg::OFF_WP, sWd = sm + Cfg::OFF_WD, sBe = sm + Cfg::OFF_BE, sBd = sm + Cfg::OFF_BD, sBp = sm + Cfg::OFF_BP;
    const uint32_t bar_w = sm + Cfg::OFF_BAR, bar_x = bar_w + 8;
    constexpr int S = Cfg::S, PW = Cfg::PW, P = Cfg::P, CEXP = Cfg::CEXP, COUT = Cfg::COUT;
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // ---- 1. loads.  TMA (one thread): expansion weights [CEXP][CIN] and projection weights [COUT][CEXP] as 16-channel
    //         slices (rows past COUT and channels past CIN are out of bounds = zero-filled), once per CTA; then -- after the
    //         PDL wait -- the input patch of the first tile (tile + halo; positions outside the image zero-filled).  The
    //         depthwise filter and the biases are loaded by the threads.
    if (tid == 0) {
        tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_we); tma_prefetch_desc(&map_wp);
        mbar_init(bar_w, 1);
        mbar_init(bar_x, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, Cfg::W_BYTES);
#pragma unroll
        for (int ks = 0; ks < Cfg::KS; ++ks) tma_load_2d(sWe + ks * Cfg::WE_SL, &map_we, bar_w, 16 * ks, 0);
        for (int ks = 0; ks < Cfg::ES; ++ks) tma_load_2d(sWp + ks * Cfg::WP_SL, &map_wp, bar_w, 16 * ks, 0);
    }
    // Expansion bias, depthwise filter and depthwise bias in the channel order of the expanded patch: position q of a
    // 16-channel slice holds channel 8 ((q >> 1) & 1) + 2 (q >> 2) + (q & 1) (see the expansion's epilogue); pairs (q even,
    // q + 1) are pairs of adjacent channels, so the fp16 filter moves as 4-byte words.
    for (int i = tid; i < 9 * CEXP / 2; i += IM_THREADS) {
        const int k = i / (CEXP / 2), w2 = i - k * (CEXP / 2), q = (2 * w2) & 15;
        const int c = ((2 * w2) & ~15) + 8 * ((q >> 1) & 1) + 2 * (q >> 2);
        reinterpret_cast<uint32_t*>(im_smem + Cfg::OFF_WD)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.wd + k * CEXP + c));
    }
    for (int i = tid; i < CEXP; i += IM_THREADS) {
        const int q = i & 15, c = (i & ~15) + 8 * ((q >> 1) & 1) + 2 * (q >> 2) + (q & 1);
        reinterpret_cast<float*>(im_smem + Cfg::OFF_BE)[i] = p.be ? __ldg(p.be + c) : 0.f;
        reinterpret_cast<float*>(im_smem + Cfg::OFF_BD)[i] = p.bd ? __ldg(p.bd + c) : 0.f;
    }
    for (int i = tid; i < Cfg::COUTP; i += IM_THREADS)
        reinterpret_cast<float*>(im_smem + Cfg::OFF_BP)[i] = (p.bp && i < COUT) ? __ldg(p.bp + i) : 0.f;
    pdl_wait();
    // tile -> (image, tile row, tile column)
    auto tile_origin = [&](int tile, int& b, int& oy0, int& ox0) {
        b = tile / p.tiles_per_img;
        const int r = tile - b * p.tiles_per_img, ty = r / p.tiles_x;
        oy0 = ty * Cfg::TH; ox0 = (r - ty * p.tiles_x) * Cfg::TW;
    };
    auto request_patch = [&](int tile) {                              // thread 0 only
        int b, oy0, ox0;
        tile_origin(tile, b, oy0, ox0);
        mbar_expect_tx(bar_x, Cfg::X_BYTES);
#pragma unroll
        for (int ks = 0; ks < Cfg::KS; ++ks)
            tma_load_4d(sIn + ks * Cfg::IN_SL, &map_x, bar_x, 16 * ks, ox0 * S - p.pad_l, oy0 * S - p.pad_t, b);
    };
    if (tid == 0 && (int)blockIdx.x < p.n_tiles) request_patch((int)blockIdx.x);
    __syncthreads();                        // barriers initialised, filter + biases visible
    mbar_wait(bar_w, 0);

    uint32_t it = 0;
    for (int tile = (int)blockIdx.x; tile < p.n_tiles; tile += (int)gridDim.x, ++it) {
    int b, oy0, ox0;
    tile_origin(tile, b, oy0, ox0);
    const int iy0 = oy0 * S - p.pad_t, ix0 = ox0 * S - p.pad_l;          // image coordinates of patch position (0, 0)
    mbar_wait(bar_x, it & 1u);              // this tile's patch has landed

    // ---- 2. expansion: units of (16-channel slice, m16 tile of positions) round-robin over the warps ----
    {
        const uint32_t six = 0x46004600u;                              // half2(6, 6)
        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = lane >> 4;         // A operand: ldmatrix lane -> (row, half)
        const int brow = (lane & 7) + (lane >> 4) * 8, bhalf = (lane >> 3) & 1;         // B operand
        // Every warp owns a contiguous range of the (slice PAIR, m16 tile) work items, slice-major: the weight fragments
        // and biases of two slices stay in registers while the warp walks over m-tiles, and one ldmatrix per k-step of the
        // positions feeds 4 KS MMAs (shared-memory wavefronts, not issue slots, bound this kernel).
        constexpr int ESP = (Cfg::ES + 1) / 2, NITEM = ESP * Cfg::MT;
        const int q0 = warp * NITEM / IM_WARPS, q1 = (warp + 1) * NITEM / IM_WARPS;
        int sp = q0 / Cfg::MT, mt = q0 - sp * Cfg::MT;
        uint32_t bq[2][Cfg::KS][4];
        float4 bias[2];
        bool fresh = true, two = true;
        for (int q = q0; q < q1; ++q) {
            if (fresh) {
                two = 2 * sp + 1 < Cfg::ES;                           // an odd slice count leaves the last pair half empty
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int jp = (h2 == 1 && !two) ? 2 * sp : 2 * sp + h2;
#pragma unroll
                    for (int ks = 0; ks < Cfg::KS; ++ks)
                        im_ldsm4(bq[h2][ks], sWe + (uint32_t)(ks * Cfg::WE_SL) + im_row(jp * 16 + brow, bhalf));
                    bias[h2] = im_lds128f(sBe + (uint32_t)(jp * 16 + 4 * t) * 4u);   // channels 4t .. 4t+3 of the slice's order
                }
                fresh = false;
            }
            uint32_t a[Cfg::KS][4];
#pragma unroll
            for (int ks = 0; ks < Cfg::KS; ++ks)
                im_ldsm4(a[ks], sIn + (uint32_t)(ks * Cfg::IN_SL) + im_row(mt * 16 + lrow, lhalf));
            // the expanded patch is LINEAR (row = 32 bytes); within a slice, position 4t + 2n + e holds channel 8n + 2t + e
            // (accumulator columns 2t, 2t+1 of n-tiles n = 0, 1): one 8-byte store per row, 8 whole rows per instruction
            const uint32_t dst0 = sMid + (uint32_t)(2 * sp * Cfg::MID_SL + (mt * 16 + g) * 32 + 8 * t);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                if (h2 == 1 && !two) break;
                float acc[2][4];
                im_mma_bias(acc[0], a[0], bq[h2][0][0], bq[h2][0][1], bias[h2].x, bias[h2].y);
                im_mma_bias(acc[1], a[0], bq[h2][0][2], bq[h2][0][3], bias[h2].z, bias[h2].w);
#pragma unroll
                for (int ks = 1; ks < Cfg::KS; ++ks) {
                    im_mma(acc[0], a[ks], bq[h2][ks][0], bq[h2][ks][1]);
                    im_mma(acc[1], a[ks], bq[h2][ks][2], bq[h2][ks][3]);
                }
                const uint32_t dst = dst0 + (uint32_t)(h2 * Cfg::MID_SL);
                im_sts64(dst, im_min2(im_cvt_relu(acc[0][0], acc[0][1]), six), im_min2(im_cvt_relu(acc[1][0], acc[1][1]), six));
                im_sts64(dst + 256, im_min2(im_cvt_relu(acc[0][2], acc[0][3]), six), im_min2(im_cvt_relu(acc[1][2], acc[1][3]), six));
            }
            if (++mt == Cfg::MT) { mt = 0; ++sp; fresh = true; }
        }
    }
    __syncthreads();
    if (tid == 0 && tile + (int)gridDim.x < p.n_tiles) {
        // every warp is done reading the patch buffer (generic proxy) -> the next tile's patch may overwrite it (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        request_patch(tile + (int)gridDim.x);
    }
    // edge tiles: positions outside the image are the depthwise layer's zero padding, not ReLU6(bias)
    if (iy0 < 0 || ix0 < 0 || iy0 + Cfg::PH > p.H || ix0 + PW > p.W) {
        for (int i = tid; i < P * Cfg::ES; i += IM_THREADS) {
            const int pos = i / Cfg::ES, sl = i - pos * Cfg::ES;
            const int py = pos / PW, px = pos - py * PW;
            if ((unsigned)(iy0 + py) >= (unsigned)p.H || (unsigned)(ix0 + px) >= (unsigned)p.W) {
                const uint32_t a = sMid + (uint32_t)(sl * Cfg::MID_SL + pos * 32);
                im_sts128(a, make_uint4(0u, 0u, 0u, 0u));
                im_sts128(a + 16, make_uint4(0u, 0u, 0u, 0u));
            }
        }
        __syncthreads();
    }

    // ---- 3. depthwise 3x3: thread = (8-channel chunk, pixel lane); packed half2 FMAs, fp16 accumulation from the fp16 bias ----
    if (tid < Cfg::NCH * Cfg::NPL) {
        const int c8 = tid % Cfg::NCH, pl = tid / Cfg::NCH;
        // The expanded patch stores a slice's 16 channels in the order of the expansion's accumulator fragments: the four
        // half2 pairs of 16-byte chunk h are channels 4h (+0,1), 8 + 4h (+0,1), 4h + 2 (+0,1), 8 + 4h + 2 (+0,1) of the slice.
        // The filter taps and the bias were stored in that order by the prologue; the outputs go back in natural order.
        const int sl = c8 >> 1, h = c8 & 1;
        uint4 w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = im_lds128(sWd + (uint32_t)(k * CEXP + 8 * c8) * 2u);
        __half2 bias4[4];
        {
            const float4 b0 = im_lds128f(sBd + (uint32_t)(8 * c8) * 4u), b1 = im_lds128f(sBd + (uint32_t)(8 * c8 + 4) * 4u);
            bias4[0] = __floats2half2_rn(b0.x, b0.y); bias4[1] = __floats2half2_rn(b0.z, b0.w);
            bias4[2] = __floats2half2_rn(b1.x, b1.y); bias4[3] = __floats2half2_rn(b1.z, b1.w);
        }
        const __half2 zero2 = __float2half2_rn(0.f), six2 = __float2half2_rn(6.f);
        const uint32_t src = sMid + (uint32_t)(sl * Cfg::MID_SL + h * 16), dstb = sDw + (uint32_t)(sl * Cfg::DW_SL + 8 * h);
        if constexpr (S == 1) {
            // stride 1: runs of 3 horizontally adjacent pixels share their input columns (15 loads per 3 outputs instead
            // of 27); columns past the tile's halo belong to discarded outputs and only read allocated shared memory
            constexpr int RPR = (Cfg::TW + 2) / 3, NRUN = Cfg::TH * RPR;
            int y = pl / RPR, xr = pl - y * RPR;
            for (int run = pl; run < NRUN; run += Cfg::NPL) {
                const int x0 = 3 * xr;
                const uint32_t a0 = src + (uint32_t)((y * PW + x0) * 32);
                __half2 acc[3][4];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) acc[i][c2] = bias4[c2];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const uint4 xv = im_lds128(a0 + (uint32_t)((ky * PW + c) * 32));
                        const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int kx = c - i;
                            if (kx >= 0 && kx < 3) {
                                const __half2* wh = reinterpret_cast<const __half2*>(&w[ky * 3 + kx]);
#pragma unroll
                                for (int c2 = 0; c2 < 4; ++c2) acc[i][c2] = __hfma2(xh[c2], wh[c2], acc[i][c2]);
                            }
                        }
                    }
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (x0 + i < Cfg::TW) {
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) oh[c2] = __hmin2(__hmax2(acc[i][c2], zero2), six2);
                        im_sts64(dstb + im_row(y * Cfg::TW + x0 + i, 0), o.x, o.z);
                        im_sts64(dstb + im_row(y * Cfg::TW + x0 + i, 1), o.y, o.w);
                    }
                xr += Cfg::NPL % RPR; y += Cfg::NPL / RPR;
                if (xr >= RPR) { xr -= RPR; ++y; }
            }
        } else {
            // stride 2: two pixels in flight per iteration (independent load -> FMA chains)
            int y = pl / Cfg::TW, x = pl - y * Cfg::TW;
            for (int px = pl; px < Cfg::NPX; px += 2 * Cfg::NPL) {
                int yy[2], xx[2];
                yy[0] = y; xx[0] = x;
                x += Cfg::NPL % Cfg::TW; y += Cfg::NPL / Cfg::TW;
                if (x >= Cfg::TW) { x -= Cfg::TW; ++y; }
                const bool two = px + Cfg::NPL < Cfg::NPX;
                yy[1] = two ? y : yy[0]; xx[1] = two ? x : xx[0];
                x += Cfg::NPL % Cfg::TW; y += Cfg::NPL / Cfg::TW;
                if (x >= Cfg::TW) { x -= Cfg::TW; ++y; }
                uint32_t a0[2];
                __half2 acc[2][4];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    a0[q] = src + (uint32_t)((yy[q] * S * PW + xx[q] * S) * 32);
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) acc[q][c2] = bias4[c2];
                }
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const __half2* wh = reinterpret_cast<const __half2*>(&w[ky * 3 + kx]);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint4 xv = im_lds128(a0[q] + (uint32_t)((ky * PW + kx) * 32));
                            const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                            for (int c2 = 0; c2 < 4; ++c2) acc[q][c2] = __hfma2(xh[c2], wh[c2], acc[q][c2]);
                        }
                    }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q == 1 && !two) break;
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) oh[c2] = __hmin2(__hmax2(acc[q][c2], zero2), six2);
                    im_sts64(dstb + im_row(px + q * Cfg::NPL, 0), o.x, o.z);
                    im_sts64(dstb + im_row(px + q * Cfg::NPL, 1), o.y, o.w);
                }
            }
        }
    }
    __syncthreads();

    // ---- 4. projection: units of (m16 tile of pixels, n-tile pair); fp32 results (+ bias) -> staging tile [pixel][COUT] ----
    {
        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = lane >> 4;
        const int brow = (lane & 7) + (lane >> 4) * 8, bhalf = (lane >> 3) & 1;
        const uint32_t sOut = sMid;
        for (int u = warp; u < Cfg::MTP * Cfg::NP; u += IM_WARPS) {
            const int mt = u / Cfg::NP, np = u - mt * Cfg::NP;
            float acc[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[n][r] = 0.f;
#pragma unroll
            for (int ks = 0; ks < Cfg::ES; ++ks) {
                uint32_t a[4], bq[4];
                im_ldsm4(a, sDw + (uint32_t)(ks * Cfg::DW_SL) + im_row(mt * 16 + lrow, lhalf));
                im_ldsm4(bq, sWp + (uint32_t)(ks * Cfg::WP_SL) + im_row(np * 16 + brow, bhalf));
                im_mma(acc[0], a, bq[0], bq[1]);
                im_mma(acc[1], a, bq[2], bq[3]);
            }
            const int r0 = mt * 16 + g;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int c = np * 16 + n * 8 + 2 * t;
                if (c < COUT) {
                    const float2 bias = im_lds64f(sBp + (uint32_t)c * 4u);
                    if (r0 < Cfg::NPX) im_sts64f(sOut + (uint32_t)(r0 * COUT + c) * 4u, acc[n][0] + bias.x, acc[n][1] + bias.y);
                    if (r0 + 8 < Cfg::NPX) im_sts64f(sOut + (uint32_t)((r0 + 8) * COUT + c) * 4u, acc[n][2] + bias.x, acc[n][3] + bias.y);
                }
            }
        }
    }
    __syncthreads();

    // ---- 5. (+ residual) -> fp16 -> 16-byte coalesced stores ----
    {
        constexpr int U = COUT / 8;
        for (int i = tid; i < Cfg::NPX * U; i += IM_THREADS) {
            const int px = i / U, u = i - px * U;
            const int y = px / Cfg::TW, x = px - y * Cfg::TW;
            const int oy = oy0 + y, ox = ox0 + x;
            if (oy < p.Ho && ox < p.Wo) {
                const float4 v0 = im_lds128f(sMid + (uint32_t)(px * COUT + 8 * u) * 4u);
                const float4 v1 = im_lds128f(sMid + (uint32_t)(px * COUT + 8 * u + 4) * 4u);
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                const size_t off = (((size_t)b * p.Ho + oy) * p.Wo + ox) * COUT + 8 * u;
                if (p.res) {
                    const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + off));
                    const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float2 f = __half22float2(rh[c]);
                        v[2 * c] += f.x; v[2 * c + 1] += f.y;
                    }
                }
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int c = 0; c < 4; ++c) oh[c] = __floats2half2_rn(v[2 * c], v[2 * c + 1]);
                *reinterpret_cast<uint4*>(p.out + off) = o;
            }
        }
    }
    __syncthreads();                        // the staging tile (= the expanded patch) is free for the next tile
    }   // tile loop
}

template <class Cfg>
static int irblock_mma_launch_t(const ssd_irblock_desc* d, cudaStream_t st) {
    ImParams p;
    p.in = static_cast<const __half*>(d->in); p.we = static_cast<const __half*>(d->exp_weight); p.be = d->exp_bias;
    p.wd = static_cast<const __half*>(d->dw_weight); p.bd = d->dw_bias;
    p.wp = static_cast<const __half*>(d->proj_weight); p.bp = d->proj_bias;
    p.res = static_cast<const __half*>(d->residual); p.out = static_cast<__half*>(d->out);
    p.B = d->B; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
    p.tiles_x = ceil_div(d->Wo, Cfg::TW);
    p.tiles_per_img = p.tiles_x * ceil_div(d->Ho, Cfg::TH);
    p.n_tiles = p.tiles_per_img * d->B;
    CUtensorMap map_x, map_we, map_wp;
    {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        uint32_t box[4] = {16, (uint32_t)Cfg::PW, (uint32_t)Cfg::PH, 1};
        int rc = cached_map(&map_x, d->in, 4, dims, str, box, nullptr, 32);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)d->Cexp};
        uint64_t str[1] = {(uint64_t)d->Cin * 2};
        uint32_t box[2] = {16, (uint32_t)Cfg::CEXP};
        int rc = cached_map(&map_we, d->exp_weight, 2, dims, str, box, nullptr, 32);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->Cexp, (uint64_t)d->Cout};
        uint64_t str[1] = {(uint64_t)d->Cexp * 2};
        uint32_t box[2] = {16, (uint32_t)Cfg::COUTP};
        int rc = cached_map(&map_wp, d->proj_weight, 2, dims, str, box, nullptr, 32);
        if (rc) return rc;
    }
    static thread_local int attr_dev = -1;
    int cur = 0;
    cudaGetDevice(&cur);
    if (attr_dev != cur) {
        cudaError_t e = cudaFuncSetAttribute(irblock_mma_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "ssd_irblock: cudaFuncSetAttribute (mma.sync variant)");
        attr_dev = cur;
    }
    const dim3 grid(min(p.n_tiles, 2 * (sm_count() > 0 ? sm_count() : 148)));
    cudaError_t le = launch_pdl(irblock_mma_kernel<Cfg>, grid, dim3(IM_THREADS), (size_t)Cfg::SMEM, st, map_x, map_we, map_wp, p);
    if (le != cudaSuccess) return cuda_fail(le, "irblock_mma_kernel");
    return SSD_OK;
}

// MobileNetV2 blocks 1-6 at width 1.0 (t = 6): (Cin, Cexp, Cout, stride) and the output tile chosen for each
typedef ImCfg<16, 96, 24, 2, 15, 5> ImBlock1;       // 150 -> 75: 5 x 15 tiles per image
typedef ImCfg<24, 144, 24, 1, 15, 5> ImBlock2;      // 75 x 75
typedef ImCfg<24, 144, 32, 2, 10, 4> ImBlock3;      // 75 -> 38
typedef ImCfg<32, 192, 32, 1, 13, 4> ImBlock45;     // 38 x 38
typedef ImCfg<32, 192, 64, 2, 10, 2> ImBlock6;      // 38 -> 19

// ------------------------------------------------------------------------------------------------------------------
// The SMALL-MAP blocks (19x19 blocks 7-12, 10x10 blocks 14-15; stride 1, Cexp = 384 ... 960): the expanded tensor of a
// tile no longer fits shared memory at once and the weights are 100-630 KB, so the expanded channels are processed in
// GROUPS of 96 (6 slices): per group  expansion -> depthwise -> projection partial sums (accumulators stay in registers
// across groups), with the group's weights streamed from L2 by TMA into a double buffer two groups ahead.  One CTA of 16
// warps per SM, one tile per CTA (4 tiles per image = 128 CTAs at batch 32).
template <int CIN_, int CEXP_, int COUT_, int TW_, int TH_, int THREADS_ = 512>
struct ImGCfg {
    static constexpr int CIN = CIN_, CEXP = CEXP_, COUT = COUT_, TW = TW_, TH = TH_;
    static constexpr int THREADS = THREADS_, WARPS = THREADS / 32;
    static constexpr int GS = 6, GCH = GS * 16, NG = CEXP / GCH;       // slices / channels per group, groups
    static constexpr int KS = CIN / 16;
    static constexpr int PW = TW + 2, PH = TH + 2, P = PW * PH, PPOS = im_up16(P);
    static constexpr int NPX = TW * TH, MPX = im_up16(NPX);
    static constexpr int COUTP = im_up16(COUT);
    static constexpr int MT = PPOS / 16, MTP = MPX / 16, NP = COUTP / 16;
    static constexpr int NUNIT = MTP * NP, UPW = (NUNIT + WARPS - 1) / WARPS;      // projection units (m-tile, n-tile pair) per warp
    static constexpr int NCH = GS * 2, NPL = THREADS / NCH;
    static constexpr int RPR = (TW + 2) / 3, NRUN = TH * RPR;
    static constexpr int IN_SL = PPOS * 32, WE_SL = GCH * 32, WP_SL = COUTP * 32;
    static constexpr int MID_SL = (PPOS + 1) * 32, DW_SL = (MPX + 1) * 32;
    static constexpr int WB_BYTES = KS * WE_SL + GS * WP_SL;           // one group's weights: [expansion | projection]
    static constexpr int OFF_IN = 0;
    static constexpr int OFF_WB = OFF_IN + KS * IN_SL;                 // [2] weight buffers
    static constexpr int OFF_MID = OFF_WB + 2 * WB_BYTES;
    static constexpr int OFF_DW = OFF_MID + (GS * MID_SL + 255) / 256 * 256;
    static constexpr int STAGE_BYTES = im_max(OFF_DW - OFF_MID + GS * DW_SL, NPX * COUT * 4);     // fp32 staging aliases [mid | dw]
    static constexpr int OFF_WD = OFF_MID + (STAGE_BYTES + 255) / 256 * 256;                      // [9][CEXP] fp16, patch channel order
    static constexpr int OFF_BE = OFF_WD + 9 * CEXP * 2;
    static constexpr int OFF_BD = OFF_BE + CEXP * 4;
    static constexpr int OFF_BP = OFF_BD + CEXP * 4;
    static constexpr int OFF_BAR = OFF_BP + COUTP * 4;                 // mbarriers: patch, weight buffer 0, weight buffer 1
    static constexpr int SMEM = OFF_BAR + 32 + 1024;
    static constexpr uint32_t X_BYTES = KS * P * 32;
    static_assert(CIN % 16 == 0 && CEXP % GCH == 0 && COUT % 8 == 0, "channel granularity");
    static_assert(OFF_WB % 256 == 0 && WB_BYTES % 256 == 0 && IN_SL % 256 == 0 && WE_SL % 256 == 0 && WP_SL % 256 == 0, "TMA destinations");
    static_assert(OFF_DW % 16 == 0 && OFF_WD % 16 == 0 && OFF_BE % 16 == 0 && OFF_BD % 16 == 0 && OFF_BP % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
    static_assert(SMEM <= 227 * 1024, "shared memory");
    static_assert(NRUN <= NPL, "one depthwise pass");
    static_assert(PW <= 256 && PH <= 256 && GCH <= 256 && COUTP <= 256, "TMA box dimensions");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
irblock_mma_grouped_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_we,
                           const __grid_constant__ CUtensorMap map_wp, const __grid_constant__ ImParams p) {
    extern __shared__ __align__(16) unsigned char im_smem_raw[];
    unsigned char* im_smem = reinterpret_cast<unsigned char*>(((uintptr_t)im_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(im_smem);
    const uint32_t sIn = sm + Cfg::OFF_IN, sWB = sm + Cfg::OFF_WB, sMid = sm + Cfg::OFF_MID, sDw = sm + Cfg::OFF_DW;
    const uint32_t sWd = sm + Cfg::OFF_WD, sBe = sm + Cfg::OFF_BE, sBd = sm + Cfg::OFF_BD, sBp = sm + Cfg::OFF_BP;
    const uint32_t bar_x = sm + Cfg::OFF_BAR, bar_w = bar_x + 8;              // bar_w + 8 * buffer
    constexpr int PW = Cfg::PW, P = Cfg::P, CEXP = Cfg::CEXP, COUT = Cfg::COUT, GS = Cfg::GS, KS = Cfg::KS;
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tile = (int)blockIdx.x;
    const int b = tile / p.tiles_per_img, tr = tile - b * p.tiles_per_img, ty = tr / p.tiles_x;
    const int oy0 = ty * Cfg::TH, ox0 = (tr - ty * p.tiles_x) * Cfg::TW;
    const int iy0 = oy0 - p.pad_t, ix0 = ox0 - p.pad_l;

    auto request_weights = [&](int grp) {                               // thread 0 only
        const uint32_t wb = sWB + (uint32_t)((grp & 1) * Cfg::WB_BYTES), bar = bar_w + 8u * (grp & 1);
        mbar_expect_tx(bar, Cfg::WB_BYTES);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) tma_load_2d(wb + ks * Cfg::WE_SL, &map_we, bar, 16 * ks, grp * Cfg::GCH);
#pragma unroll
        for (int j = 0; j < GS; ++j) tma_load_2d(wb + KS * Cfg::WE_SL + j * Cfg::WP_SL, &map_wp, bar, 16 * (grp * GS + j), 0);
    };
    if (tid == 0) {
        tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_we); tma_prefetch_desc(&map_wp);
        mbar_init(bar_x, 1); mbar_init(bar_w, 1); mbar_init(bar_w + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        request_weights(0);
        if (Cfg::NG > 1) request_weights(1);
    }
    // depthwise filter and biases in the channel order of the expanded patch (see irblock_mma_kernel)
    for (int i = tid; i < 9 * CEXP / 2; i += Cfg::THREADS) {
        const int k = i / (CEXP / 2), w2 = i - k * (CEXP / 2), q = (2 * w2) & 15;
        const int c = ((2 * w2) & ~15) + 8 * ((q >> 1) & 1) + 2 * (q >> 2);
        reinterpret_cast<uint32_t*>(im_smem + Cfg::OFF_WD)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.wd + k * CEXP + c));
    }
    for (int i = tid; i < CEXP; i += Cfg::THREADS) {
        const int q = i & 15, c = (i & ~15) + 8 * ((q >> 1) & 1) + 2 * (q >> 2) + (q & 1);
        reinterpret_cast<float*>(im_smem + Cfg::OFF_BE)[i] = p.be ? __ldg(p.be + c) : 0.f;
        reinterpret_cast<float*>(im_smem + Cfg::OFF_BD)[i] = p.bd ? __ldg(p.bd + c) : 0.f;
    }
    for (int i = tid; i < Cfg::COUTP; i += Cfg::THREADS)
        reinterpret_cast<float*>(im_smem + Cfg::OFF_BP)[i] = (p.bp && i < COUT) ? __ldg(p.bp + i) : 0.f;
    pdl_wait();
    if (tid == 0) {
        mbar_expect_tx(bar_x, Cfg::X_BYTES);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) tma_load_4d(sIn + ks * Cfg::IN_SL, &map_x, bar_x, 16 * ks, ix0, iy0, b);
    }
    __syncthreads();
    mbar_wait(bar_x, 0);

    const uint32_t six = 0x46004600u;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = lane >> 4;       // A operand: ldmatrix lane -> (row, half)
    const int brow = (lane & 7) + (lane >> 4) * 8, bhalf = (lane >> 3) & 1;       // B operand
    const bool edge = iy0 < 0 || ix0 < 0 || iy0 + Cfg::PH > p.H || ix0 + PW > p.W;
    float pacc[Cfg::UPW][2][4];                                                   // projection partial sums of this warp's units
#pragma unroll
    for (int i = 0; i < Cfg::UPW; ++i)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) pacc[i][n][r] = 0.f;

    for (int grp = 0; grp < Cfg::NG; ++grp) {
        const uint32_t sWe = sWB + (uint32_t)((grp & 1) * Cfg::WB_BYTES), sWp = sWe + KS * Cfg::WE_SL;
        mbar_wait(bar_w + 8u * (grp & 1), (uint32_t)(grp >> 1) & 1u);
        // ---- expansion of this group's 6 slices: contiguous (slice, m-tile) pairs per warp ----
        {
            constexpr int NPAIR = GS * Cfg::MT;
            const int q0 = warp * NPAIR / Cfg::WARPS, q1 = (warp + 1) * NPAIR / Cfg::WARPS;
            int jp = q0 / Cfg::MT, mt = q0 - jp * Cfg::MT;
            uint32_t bq[KS][4];
            float4 bias;
            bool fresh = true;
            for (int q = q0; q < q1; ++q) {
                if (fresh) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                        im_ldsm4(bq[ks], sWe + (uint32_t)(ks * Cfg::WE_SL) + im_row(jp * 16 + brow, bhalf));
                    bias = im_lds128f(sBe + (uint32_t)(grp * Cfg::GCH + jp * 16 + 4 * t) * 4u);
                    fresh = false;
                }
                float acc[2][4];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t a[4];
                    im_ldsm4(a, sIn + (uint32_t)(ks * Cfg::IN_SL) + im_row(mt * 16 + lrow, lhalf));
                    if (ks == 0) {
                        im_mma_bias(acc[0], a, bq[0][0], bq[0][1], bias.x, bias.y);
                        im_mma_bias(acc[1], a, bq[0][2], bq[0][3], bias.z, bias.w);
                    } else {
                        im_mma(acc[0], a, bq[ks][0], bq[ks][1]);
                        im_mma(acc[1], a, bq[ks][2], bq[ks][3]);
                    }
                }
                const uint32_t dst = sMid + (uint32_t)(jp * Cfg::MID_SL + (mt * 16 + g) * 32 + 8 * t);
                im_sts64(dst, im_min2(im_cvt_relu(acc[0][0], acc[0][1]), six), im_min2(im_cvt_relu(acc[1][0], acc[1][1]), six));
                im_sts64(dst + 256, im_min2(im_cvt_relu(acc[0][2], acc[0][3]), six), im_min2(im_cvt_relu(acc[1][2], acc[1][3]), six));
                if (++mt == Cfg::MT) { mt = 0; ++jp; fresh = true; }
            }
        }
        __syncthreads();
        if (edge) {                                   // CTA-uniform: zero the positions outside the image (depthwise padding)
            for (int i = tid; i < P * GS; i += Cfg::THREADS) {
                const int pos = i / GS, sl = i - pos * GS;
                const int py = pos / PW, px = pos - py * PW;
                if ((unsigned)(iy0 + py) >= (unsigned)p.H || (unsigned)(ix0 + px) >= (unsigned)p.W) {
                    const uint32_t a = sMid + (uint32_t)(sl * Cfg::MID_SL + pos * 32);
                    im_sts128(a, make_uint4(0u, 0u, 0u, 0u));
                    im_sts128(a + 16, make_uint4(0u, 0u, 0u, 0u));
                }
            }
            __syncthreads();
        }
        // ---- depthwise 3x3 (stride 1): thread = (8-channel chunk, run of 3 adjacent pixels) ----
        if (tid < Cfg::NCH * Cfg::NPL) {
            const int c8 = tid % Cfg::NCH, pl = tid / Cfg::NCH;
            if (pl < Cfg::NRUN) {
                const int sl = c8 >> 1, h = c8 & 1;
                const int y = pl / Cfg::RPR, x0 = 3 * (pl - y * Cfg::RPR);
                const uint32_t wsrc = sWd + (uint32_t)(grp * Cfg::GCH + 8 * c8) * 2u;
                __half2 bias4[4];
                {
                    const float4 b0 = im_lds128f(sBd + (uint32_t)(grp * Cfg::GCH + 8 * c8) * 4u);
                    const float4 b1 = im_lds128f(sBd + (uint32_t)(grp * Cfg::GCH + 8 * c8 + 4) * 4u);
                    bias4[0] = __floats2half2_rn(b0.x, b0.y); bias4[1] = __floats2half2_rn(b0.z, b0.w);
                    bias4[2] = __floats2half2_rn(b1.x, b1.y); bias4[3] = __floats2half2_rn(b1.z, b1.w);
                }
                const uint32_t a0 = sMid + (uint32_t)(sl * Cfg::MID_SL + h * 16 + (y * PW + x0) * 32);
                __half2 acc[3][4];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) acc[i][c2] = bias4[c2];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    uint4 w3[3];
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) w3[kx] = im_lds128(wsrc + (uint32_t)((ky * 3 + kx) * CEXP) * 2u);
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const uint4 xv = im_lds128(a0 + (uint32_t)((ky * PW + c) * 32));
                        const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int kx = c - i;
                            if (kx >= 0 && kx < 3) {
                                const __half2* wh = reinterpret_cast<const __half2*>(&w3[kx]);
#pragma unroll
                                for (int c2 = 0; c2 < 4; ++c2) acc[i][c2] = __hfma2(xh[c2], wh[c2], acc[i][c2]);
                            }
                        }
                    }
                }
                const __half2 zero2 = __float2half2_rn(0.f), six2 = __float2half2_rn(6.f);
                const uint32_t dstb = sDw + (uint32_t)(sl * Cfg::DW_SL + 8 * h);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (x0 + i < Cfg::TW) {
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) oh[c2] = __hmin2(__hmax2(acc[i][c2], zero2), six2);
                        im_sts64(dstb + im_row(y * Cfg::TW + x0 + i, 0), o.x, o.z);
                        im_sts64(dstb + im_row(y * Cfg::TW + x0 + i, 1), o.y, o.w);
                    }
            }
        }
        __syncthreads();
        // ---- projection partial sums: this group's 6 k16 steps into the warp's units ----
#pragma unroll
        for (int i = 0; i < Cfg::UPW; ++i) {
            const int u = warp + i * Cfg::WARPS;
            if (u < Cfg::NUNIT) {
                const int mt = u / Cfg::NP, np = u - mt * Cfg::NP;
#pragma unroll
                for (int ks = 0; ks < GS; ++ks) {
                    uint32_t a[4], bq[4];
                    im_ldsm4(a, sDw + (uint32_t)(ks * Cfg::DW_SL) + im_row(mt * 16 + lrow, lhalf));
                    im_ldsm4(bq, sWp + (uint32_t)(ks * Cfg::WP_SL) + im_row(np * 16 + brow, bhalf));
                    im_mma(pacc[i][0], a, bq[0], bq[1]);
                    im_mma(pacc[i][1], a, bq[2], bq[3]);
                }
            }
        }
        __syncthreads();                              // weight buffer, expanded patch and depthwise tile of this group are free
        if (tid == 0 && grp + 2 < Cfg::NG) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            request_weights(grp + 2);
        }
    }

    // ---- projection results (+ bias) -> fp32 staging tile [pixel][COUT] (aliases the expanded patch / depthwise tile) ----
    {
        const uint32_t sOut = sMid;
#pragma unroll
        for (int i = 0; i < Cfg::UPW; ++i) {
            const int u = warp + i * Cfg::WARPS;
            if (u < Cfg::NUNIT) {
                const int mt = u / Cfg::NP, np = u - mt * Cfg::NP;
                const int r0 = mt * 16 + g;
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    const int c = np * 16 + n * 8 + 2 * t;
                    if (c < COUT) {
                        const float2 bias = im_lds64f(sBp + (uint32_t)c * 4u);
                        if (r0 < Cfg::NPX) im_sts64f(sOut + (uint32_t)(r0 * COUT + c) * 4u, pacc[i][n][0] + bias.x, pacc[i][n][1] + bias.y);
                        if (r0 + 8 < Cfg::NPX) im_sts64f(sOut + (uint32_t)((r0 + 8) * COUT + c) * 4u, pacc[i][n][2] + bias.x, pacc[i][n][3] + bias.y);
                    }
                }
            }
        }
    }
    __syncthreads();
    {
        constexpr int U = COUT / 8;
        for (int i = tid; i < Cfg::NPX * U; i += Cfg::THREADS) {
            const int px = i / U, u = i - px * U;
            const int y = px / Cfg::TW, x = px - y * Cfg::TW;
            const int oy = oy0 + y, ox = ox0 + x;
            if (oy < p.Ho && ox < p.Wo) {
                const float4 v0 = im_lds128f(sMid + (uint32_t)(px * COUT + 8 * u) * 4u);
                const float4 v1 = im_lds128f(sMid + (uint32_t)(px * COUT + 8 * u + 4) * 4u);
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                const size_t off = (((size_t)b * p.Ho + oy) * p.Wo + ox) * COUT + 8 * u;
                if (p.res) {
                    const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + off));
                    const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float2 f = __half22float2(rh[c]);
                        v[2 * c] += f.x; v[2 * c + 1] += f.y;
                    }
                }
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int c = 0; c < 4; ++c) oh[c] = __floats2half2_rn(v[2 * c], v[2 * c + 1]);
                *reinterpret_cast<uint4*>(p.out + off) = o;
            }
        }
    }
}

template <class Cfg>
static int irblock_mma_grouped_launch_t(const ssd_irblock_desc* d, cudaStream_t st) {
    ImParams p;
    p.in = static_cast<const __half*>(d->in); p.we = static_cast<const __half*>(d->exp_weight); p.be = d->exp_bias;
    p.wd = static_cast<const __half*>(d->dw_weight); p.bd = d->dw_bias;
    p.wp = static_cast<const __half*>(d->proj_weight); p.bp = d->proj_bias;
    p.res = static_cast<const __half*>(d->residual); p.out = static_cast<__half*>(d->out);
    p.B = d->B; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
    p.tiles_x = ceil_div(d->Wo, Cfg::TW);
    p.tiles_per_img = p.tiles_x * ceil_div(d->Ho, Cfg::TH);
    p.n_tiles = p.tiles_per_img * d->B;
    CUtensorMap map_x, map_we, map_wp;
    {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        uint32_t box[4] = {16, (uint32_t)Cfg::PW, (uint32_t)Cfg::PH, 1};
        int rc = cached_map(&map_x, d->in, 4, dims, str, box, nullptr, 32);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)d->Cexp};
        uint64_t str[1] = {(uint64_t)d->Cin * 2};
        uint32_t box[2] = {16, (uint32_t)Cfg::GCH};
        int rc = cached_map(&map_we, d->exp_weight, 2, dims, str, box, nullptr, 32);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->Cexp, (uint64_t)d->Cout};
        uint64_t str[1] = {(uint64_t)d->Cexp * 2};
        uint32_t box[2] = {16, (uint32_t)Cfg::COUTP};
        int rc = cached_map(&map_wp, d->proj_weight, 2, dims, str, box, nullptr, 32);
        if (rc) return rc;
    }
    static thread_local int attr_dev = -1;
    int cur = 0;
    cudaGetDevice(&cur);
    if (attr_dev != cur) {
        cudaError_t e = cudaFuncSetAttribute(irblock_mma_grouped_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "ssd_irblock: cudaFuncSetAttribute (grouped mma.sync variant)");
        attr_dev = cur;
    }
    cudaError_t le = launch_pdl(irblock_mma_grouped_kernel<Cfg>, dim3(p.n_tiles), dim3(Cfg::THREADS), (size_t)Cfg::SMEM, st,
                                map_x, map_we, map_wp, p);
    if (le != cudaSuccess) return cuda_fail(le, "irblock_mma_grouped_kernel");
    return SSD_OK;
}

// MobileNetV2 blocks 7-12 (19 x 19: 4 tiles of 19 x 5 per image) and 14-15 (10 x 10: 4 tiles of 10 x 3)
typedef ImGCfg<64, 384, 64, 19, 5> ImBlock7;
typedef ImGCfg<64, 384, 96, 19, 5> ImBlock10;
typedef ImGCfg<96, 576, 96, 19, 5> ImBlock11;
typedef ImGCfg<160, 960, 160, 10, 3, 768> ImBlock14;      // 24 warps: 39 vs 41 us (the 19 x 19 configurations lose with more warps)

// -1 automatic (large-map variant always; the channel-grouped small-map variant when programmatic dependent launch is on:
// its weight prologue then overlaps the predecessor, which is what makes it faster than the tcgen05 kernel), 0 tcgen05 kernel only, 1 every mma.sync variant that has an instantiation,
// 2 the large-map mma.sync variant only (blocks 1-6); SSD_B200_IRBLOCK presets it
static int irblock_mode_from_env() {
    const char* e = getenv("SSD_B200_IRBLOCK");
    return e ? atoi(e) : -1;
}
static int g_irblock_mode = irblock_mode_from_env();

// 1 when the mma.sync variant has an instantiation for this block (and ReLU6 / ReLU6 / linear activations, B <= 65535)
bool conv_irblock_mma_matches(const ssd_irblock_desc* d) {
    if (g_irblock_mode == 0) return false;
    if (!(d->exp_act == SSD_ACT_RELU6 && d->dw_act == SSD_ACT_RELU6 && d->act == SSD_ACT_NONE && d->B <= 65535)) return false;
    auto is = [&](int ci, int ce, int co, int s) { return d->Cin == ci && d->Cexp == ce && d->Cout == co && d->stride == s; };
    const bool large = is(16, 96, 24, 2) || is(24, 144, 24, 1) || is(24, 144, 32, 2) || is(32, 192, 32, 1) || is(32, 192, 64, 2);
    const bool small = is(64, 384, 64, 1) || is(64, 384, 96, 1) || is(96, 576, 96, 1) || is(160, 960, 160, 1);
    return large || (small && (g_irblock_mode == 1 || (g_irblock_mode == -1 && pdl_enabled())));
}

int conv_irblock_mma_launch(const ssd_irblock_desc* d, cudaStream_t st) {
    if (d->Cexp == 384) return d->Cout == 64 ? irblock_mma_grouped_launch_t<ImBlock7>(d, st) : irblock_mma_grouped_launch_t<ImBlock10>(d, st);
    if (d->Cexp == 576) return irblock_mma_grouped_launch_t<ImBlock11>(d, st);
    if (d->Cexp == 960) return irblock_mma_grouped_launch_t<ImBlock14>(d, st);
    if (d->Cexp == 96) return irblock_mma_launch_t<ImBlock1>(d, st);
    if (d->Cexp == 144) return d->stride == 1 ? irblock_mma_launch_t<ImBlock2>(d, st) : irblock_mma_launch_t<ImBlock3>(d, st);
    return d->stride == 1 ? irblock_mma_launch_t<ImBlock45>(d, st) : irblock_mma_launch_t<ImBlock6>(d, st);
}

}  // namespace ssd

extern "C" int ssd_debug_irblock_mode(int mode) {
    ssd::g_irblock_mode = mode;
    return SSD_OK;
}
