// Implicit-GEMM convolution for sm_100a (baseline tensor-core path: legacy
// mma.sync HMMA with a cp.async multi-stage pipeline).  Handles every Conv2D of
// the two SSD graphs: 1x1 / 3x3, stride 1/2, asymmetric TensorFlow padding,
// dilation 6 (conv6), fused bias + ReLU/ReLU6 + residual add, fp16 or fp32
// output, and a two-segment strided output so the multibox head writes
// straight into the concatenated [B,N,L] / [B,N,4] tensors.
//
//   GEMM view:  out[m, n] = sum_{tap, c} in[pix(m, tap), c] * w[n, tap, c]
//   M = B*Ho*Wo (NHWC rows), N = Cout, K = KH*KW*Cin.
//
// The tcgen05/TMEM path for the GEMM-shaped layers lives in gemm_tcgen05.cu;
// this kernel is the general fallback and the numerical cross-check for it.

#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace ssd {

constexpr int BK = 32;            // k-chunk: 32 halfs = 64 B per row = 4 x 16-byte pieces
constexpr int STAGES = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 16 : 0;       // src-size 0 -> 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte piece `p` (0..3) of row `r` inside a [rows][BK] half tile (XOR swizzle)
__device__ __forceinline__ uint32_t tile_off(int r, int p) { return (uint32_t)(r * 64 + ((p ^ ((r >> 1) & 3)) << 4)); }

struct ConvK {                     // kernel-side copy of ssd_conv_desc
    const __half* in; const __half* w; const float* bias; const __half* res;
    void* out0; void* out1;
    int B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, dil, pad_t, pad_l, act, out_f32, split;
    long long img0, pix0, img1, pix1;
    int M, HoWo, chunks_per_tap, n_chunks;
};

template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32)
conv_igemm_kernel(const ConvK p) {
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int MT = WM / 16, NTL = WN / 8;
    static_assert(WM % 16 == 0 && WN % 16 == 0, "warp tile must be a multiple of 16x16");
    constexpr int A_STAGE = BM * BK * 2, B_STAGE = BN * BK * 2;
    constexpr int A_PIECES = BM * 4, B_PIECES = BN * 4;
    constexpr int A_ITERS = (A_PIECES + NT - 1) / NT, B_ITERS = (B_PIECES + NT - 1) / NT;
    constexpr int OUT_PITCH = BN + 4;                      // fp32 staging pitch

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + STAGES * A_STAGE;
    float* sOut = reinterpret_cast<float*>(smem);          // reused after the main loop

    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // ---- per-thread gather state for the A (im2col) pieces it owns ----------
    int a_row[A_ITERS], a_piece[A_ITERS], a_iy0[A_ITERS], a_ix0[A_ITERS];
    const __half* a_base[A_ITERS];
    bool a_ok[A_ITERS];
#pragma unroll
    for (int i = 0; i < A_ITERS; ++i) {
        int q = tid + i * NT;
        int r = q >> 2;
        a_row[i] = r; a_piece[i] = q & 3;
        int m = m0 + r;
        a_ok[i] = (q < A_PIECES) && (m < p.M);
        int mm = a_ok[i] ? m : 0;
        int b = mm / p.HoWo, pix = mm - b * p.HoWo;
        int oy = pix / p.Wo, ox = pix - oy * p.Wo;
        a_iy0[i] = oy * p.stride - p.pad_t;
        a_ix0[i] = ox * p.stride - p.pad_l;
        a_base[i] = p.in + (size_t)b * p.H * p.W * p.Cin;
    }
    const int Kw = p.KH * p.KW * p.Cin;                    // weight row length (halfs)

    auto load_chunk = [&](int chunk, int stage) {
        int tap = chunk / p.chunks_per_tap;
        int c0 = (chunk - tap * p.chunks_per_tap) * BK;
        int ky = tap / p.KW, kx = tap - ky * p.KW;
        uint32_t a_dst = smem_u32(sA + stage * A_STAGE), b_dst = smem_u32(sB + stage * B_STAGE);
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            if (A_PIECES % NT != 0 && tid + i * NT >= A_PIECES) break;
            int iy = a_iy0[i] + ky * p.dil, ix = a_ix0[i] + kx * p.dil;
            int c = c0 + a_piece[i] * 8;
            bool v = a_ok[i] && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W && c < p.Cin;
            const __half* src = v ? a_base[i] + ((size_t)iy * p.W + ix) * p.Cin + c : p.in;
            cp_async16(a_dst + tile_off(a_row[i], a_piece[i]), src, v);
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            int q = tid + i * NT;
            if (B_PIECES % NT != 0 && q >= B_PIECES) break;
            int r = q >> 2, pc = q & 3;
            int n = n0 + r, c = c0 + pc * 8;
            bool v = n < p.Cout && c < p.Cin;
            const __half* src = v ? p.w + (size_t)n * Kw + (size_t)tap * p.Cin + c : p.w;
            cp_async16(b_dst + tile_off(r, pc), src, v);
        }
    };

    float acc[MT][NTL][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

    // ---- software pipeline -----------------------------------------------------
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < p.n_chunks) load_chunk(s, s);
        cp_async_commit();
    }
    for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        int nxt = chunk + STAGES - 1;
        if (nxt < p.n_chunks) load_chunk(nxt, nxt % STAGES);
        cp_async_commit();

        const int stage = chunk % STAGES;
        const uint32_t a_s = smem_u32(sA + stage * A_STAGE), b_s = smem_u32(sB + stage * B_STAGE);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[MT][4], bf[NTL / 2][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                int r = warp_m * WM + mt * 16 + (lane & 15);
                ldmatrix_x4(af[mt], a_s + tile_off(r, ks * 2 + (lane >> 4)));
            }
#pragma unroll
            for (int np = 0; np < NTL / 2; ++np) {
                int r = warp_n * WN + np * 16 + ((lane >> 4) << 3) + (lane & 7);
                ldmatrix_x4(bf[np], b_s + tile_off(r, ks * 2 + ((lane >> 3) & 1)));
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt)
                    mma16816(acc[mt][nt], af[mt], bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                       // pipeline smem is free: reuse for staging

    // ---- epilogue 1: bias + activation, fragments -> fp32 smem tile -----------
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            int col = warp_n * WN + nt * 8 + (lane & 3) * 2;
            int n = n0 + col;
            float b0 = (p.bias && n < p.Cout) ? __ldg(p.bias + n) : 0.0f;
            float b1 = (p.bias && n + 1 < p.Cout) ? __ldg(p.bias + n + 1) : 0.0f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int row = warp_m * WM + mt * 16 + (lane >> 2) + h * 8;
                float v0 = acc[mt][nt][h * 2] + b0, v1 = acc[mt][nt][h * 2 + 1] + b1;
                if (p.act == SSD_ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                else if (p.act == SSD_ACT_RELU6) { v0 = fminf(fmaxf(v0, 0.f), 6.f); v1 = fminf(fmaxf(v1, 0.f), 6.f); }
                *reinterpret_cast<float2*>(sOut + row * OUT_PITCH + col) = make_float2(v0, v1);
            }
        }
    __syncthreads();

    // ---- epilogue 2: coalesced stores (+ residual), 8 channels per thread ------
    const bool single = p.split >= p.Cout;
    const bool vec_ok = !p.out_f32 && single && (p.pix0 % 8 == 0) && (p.img0 % 8 == 0) && (p.Cout % 8 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out0) & 15) == 0) &&
                        (p.res == nullptr || (reinterpret_cast<uintptr_t>(p.res) & 15) == 0);
    if (vec_ok) {
        constexpr int GROUPS = BN / 8;
        for (int e = tid; e < BM * GROUPS; e += NT) {
            int row = e / GROUPS, g = e - row * GROUPS;
            int m = m0 + row, n = n0 + g * 8;
            if (m >= p.M || n >= p.Cout) continue;
            int b = m / p.HoWo, pix = m - b * p.HoWo;
            size_t off = (size_t)b * p.img0 + (size_t)pix * p.pix0 + n;
            const float* s = sOut + row * OUT_PITCH + g * 8;
            float4 lo = *reinterpret_cast<const float4*>(s), hi = *reinterpret_cast<const float4*>(s + 4);
            float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            if (p.res) {
                uint4 r = __ldg(reinterpret_cast<const uint4*>(p.res + off));
                const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float2 f = __half22float2(rh[k]);
                    v[2 * k] += f.x; v[2 * k + 1] += f.y;
                }
            }
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out0) + off) = o;
        }
    } else {
        for (int e = tid; e < BM * BN; e += NT) {
            int row = e / BN, col = e - row * BN;
            int m = m0 + row, n = n0 + col;
            if (m >= p.M || n >= p.Cout) continue;
            int b = m / p.HoWo, pix = m - b * p.HoWo;
            float v = sOut[row * OUT_PITCH + col];
            const bool seg1 = n >= p.split;
            size_t off = seg1 ? (size_t)b * p.img1 + (size_t)pix * p.pix1 + (n - p.split)
                              : (size_t)b * p.img0 + (size_t)pix * p.pix0 + n;
            void* base = seg1 ? p.out1 : p.out0;
            if (p.res && !seg1) v += __half2float(p.res[off]);
            if (p.out_f32) reinterpret_cast<float*>(base)[off] = v;
            else reinterpret_cast<__half*>(base)[off] = __float2half_rn(v);
        }
    }
}

template <int BM, int BN, int WARPS_M, int WARPS_N>
static int launch_conv(const ConvK& k, cudaStream_t st) {
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr size_t pipe = (size_t)STAGES * (BM + BN) * BK * 2;
    constexpr size_t stage = (size_t)BM * (BN + 4) * 4;
    constexpr size_t smem = pipe > stage ? pipe : stage;
    auto kern = conv_igemm_kernel<BM, BN, WARPS_M, WARPS_N>;
    static bool configured = false;      // per-process; attribute is per-function (device-agnostic enough here)
    if (!configured || true) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "conv_igemm: cudaFuncSetAttribute");
        configured = true;
    }
    dim3 grid(ceil_div(k.M, BM), ceil_div(k.Cout, BN));
    cudaError_t le = launch_pdl(kern, grid, dim3(NT), smem, st, k);
    if (le != cudaSuccess) return cuda_fail(le, "conv_igemm_kernel");
    return SSD_OK;
}

// conv_tcgen05.cu
bool conv_tcgen05_supported(const ssd_conv_desc* d);
bool conv_dwproj_supported(const ssd_dwproj_desc* d);
int  conv_dwproj_launch(const ssd_dwproj_desc* d, cudaStream_t st);
int  conv_tcgen05_launch(const ssd_conv_desc* d, cudaStream_t st);

// SSD_B200_CONV=legacy forces the mma.sync kernel for every convolution (A/B measurements, cross-checks).
static bool force_legacy_conv() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SSD_B200_CONV");
        v = (e && strcmp(e, "legacy") == 0) ? 1 : 0;
    }
    return v == 1;
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_conv2d(const ssd_conv_desc* d, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d);
    SSD_REQUIRE_PTR(d->in); SSD_REQUIRE_PTR(d->weight); SSD_REQUIRE_PTR(d->out0);
    SSD_REQUIRE(d->B >= 1 && d->H >= 1 && d->W >= 1 && d->Cin >= 8 && d->Cin % 8 == 0 && d->Cout >= 1 &&
                d->Ho >= 1 && d->Wo >= 1 && d->KH >= 1 && d->KW >= 1 && d->KH * d->KW <= 49 &&
                d->stride >= 1 && d->dilation >= 1, SSD_ERR_SHAPE,
                "ssd_conv2d: bad shape B=%d H=%d W=%d Cin=%d Cout=%d Ho=%d Wo=%d k=%dx%d s=%d d=%d (Cin must be a multiple of 8)",
                d->B, d->H, d->W, d->Cin, d->Cout, d->Ho, d->Wo, d->KH, d->KW, d->stride, d->dilation);
    SSD_REQUIRE((int64_t)d->B * d->Ho * d->Wo < ((int64_t)1 << 31), SSD_ERR_SHAPE, "ssd_conv2d: M overflows int32");
    SSD_REQUIRE(d->split >= 0 && d->split <= d->Cout, SSD_ERR_SHAPE, "ssd_conv2d: split=%d outside [0,%d]", d->split, d->Cout);
    if (d->split < d->Cout) SSD_REQUIRE_PTR(d->out1);
    SSD_REQUIRE(d->act >= SSD_ACT_NONE && d->act <= SSD_ACT_RELU6, SSD_ERR_SHAPE, "ssd_conv2d: bad activation %d", d->act);
    ConvK k;
    k.in = (const __half*)d->in; k.w = (const __half*)d->weight; k.bias = d->bias; k.res = (const __half*)d->residual;
    k.out0 = d->out0; k.out1 = d->out1;
    k.B = d->B; k.H = d->H; k.W = d->W; k.Cin = d->Cin; k.Ho = d->Ho; k.Wo = d->Wo; k.Cout = d->Cout;
    k.KH = d->KH; k.KW = d->KW; k.stride = d->stride; k.dil = d->dilation; k.pad_t = d->pad_top; k.pad_l = d->pad_left;
    k.act = d->act; k.out_f32 = d->out_f32; k.split = d->split;
    k.img0 = d->img_stride0; k.pix0 = d->pix_stride0; k.img1 = d->img_stride1; k.pix1 = d->pix_stride1;
    k.M = d->B * d->Ho * d->Wo; k.HoWo = d->Ho * d->Wo;
    k.chunks_per_tap = (d->Cin + BK - 1) / BK;
    k.n_chunks = k.chunks_per_tap * d->KH * d->KW;
    cudaStream_t st = as_stream(stream);
    if (!force_legacy_conv() && conv_tcgen05_supported(d)) return conv_tcgen05_launch(d, st);

    // Tile choice: narrow N tiles for the thin MobileNetV2 projections, smaller M
    // tiles when the grid would not cover the 148 SMs.
    const int sms = sm_count();
    const int N = d->Cout;
    if (N <= 32) {
        return launch_conv<128, 32, 4, 2>(k, st);
    } else if (N <= 64) {
        if (ceil_div(k.M, 128) >= sms) return launch_conv<128, 64, 4, 2>(k, st);
        return launch_conv<64, 64, 2, 4>(k, st);
    } else {
        long tiles128 = (long)ceil_div(k.M, 128) * ceil_div(N, 128);
        if (tiles128 >= 2L * sms || N > 96 && tiles128 >= sms) return launch_conv<128, 128, 2, 4>(k, st);
        long tiles64 = (long)ceil_div(k.M, 128) * ceil_div(N, 64);
        if (tiles64 >= sms) return launch_conv<128, 64, 4, 2>(k, st);
        return launch_conv<64, 64, 2, 4>(k, st);
    }
}

extern "C" int ssd_dwproj(const ssd_dwproj_desc* d, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d);
    SSD_REQUIRE_PTR(d->in); SSD_REQUIRE_PTR(d->dw_weight); SSD_REQUIRE_PTR(d->proj_weight); SSD_REQUIRE_PTR(d->out);
    SSD_REQUIRE(d->B >= 1 && d->H >= 1 && d->W >= 1 && d->C >= 8 && d->Ho >= 1 && d->Wo >= 1 && d->Cout >= 8 &&
                d->dw_act >= SSD_ACT_NONE && d->dw_act <= SSD_ACT_RELU6 && d->act >= SSD_ACT_NONE && d->act <= SSD_ACT_RELU6,
                SSD_ERR_SHAPE, "ssd_dwproj: bad shape B=%d H=%d W=%d C=%d Ho=%d Wo=%d Cout=%d", d->B, d->H, d->W, d->C, d->Ho,
                d->Wo, d->Cout);
    SSD_REQUIRE(ssd::conv_dwproj_supported(d), SSD_ERR_UNSUPPORTED,
                "ssd_dwproj: unsupported configuration (C %% 8, Cout %% 8, Cout <= 256, stride 1|2, 16-byte aligned pointers)");
    return ssd::conv_dwproj_launch(d, as_stream(stream));
}

extern "C" int ssd_dwproj_supported(const ssd_dwproj_desc* d) {
    if (d == nullptr || d->in == nullptr || d->dw_weight == nullptr || d->proj_weight == nullptr || d->out == nullptr) return 0;
    return ssd::conv_dwproj_supported(d) ? 1 : 0;
}
