// Training-side kernels of the tf-ssd hot path (trainer.py:86-127: Keras fit =
// forward + loss + backward + Adam) for sm_100a.
//
//   * conv_wgrad_kernel      filter gradient as a split-K implicit GEMM on the tensor
//                            cores: dW[co, tap, ci] = sum_pixels dY[pix, co] * X[pix+tap, ci]
//                            (both operands are pixel-major in memory, so fragments come
//                            from ldmatrix.trans; fp32 atomics combine the K splits)
//   * data gradient          is NOT here: for stride 1 it is the forward convolution of dY
//                            with the flipped / transposed filter, so it runs on the
//                            tcgen05 kernel (conv_tcgen05.cu); stride 2 zero-upsamples dY first
//   * element-wise / reduction helpers: ReLU mask, bias gradient, filter flip-transpose,
//     zero-upsample, max-pool backward, L2Normalization backward, head gradient gather,
//     fused Adam (fp32 master + fp16 working copy, L2 regulariser, loss-scale removal).

#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace ssd {

__device__ __forceinline__ uint32_t smem_u32t(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16t(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816t(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------ wgrad --
constexpr int WG_BM = 128, WG_BN = 128, WG_BK = 32, WG_STAGES = 4, WG_THREADS = 256;

struct WgradK {
    const __half* x; const __half* dy; float* dw;
    int B, H, W, Cin, Ho, Wo, Cout, KW, stride, dil, pad_t, pad_l;
    int M, HoWo, ldy, taps;
    int chunks_total, chunks_per_split, tiles_n;
};

// byte offset of 16-byte piece p (0..15) of pixel row k inside a [32][128] half tile (XOR swizzle)
__device__ __forceinline__ uint32_t wg_off(int k, int p) { return (uint32_t)(k * 256 + ((p ^ (k & 7)) << 4)); }

__global__ void __launch_bounds__(WG_THREADS)
conv_wgrad_kernel(const WgradK p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int A_STAGE = WG_BK * WG_BM * 2, B_STAGE = WG_BK * WG_BN * 2;
    unsigned char* sA = smem;                              // dY tile  [pix][co]
    unsigned char* sB = smem + WG_STAGES * A_STAGE;        // X tile   [pix][ci]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp & 1, warp_n = warp >> 1;       // 2 x 4 warps, warp tile 64 (co) x 32 (ci)
    // The GEMM N dimension is the merged (tap, ci) index n = tap*Cin + ci -- dW[co][tap][ci] is contiguous in n --
    // so a dY tile is read once per 128 merged columns instead of once per tap (Cin = 8: one tile for all 9 taps).
    const int tile_m = blockIdx.x / p.tiles_n, tile_n = blockIdx.x - tile_m * p.tiles_n;
    const int co0 = tile_m * WG_BM, n0 = tile_n * WG_BN;
    const int NT = p.taps * p.Cin;
    const int c_begin = blockIdx.z * p.chunks_per_split;
    const int c_end = min(p.chunks_total, c_begin + p.chunks_per_split);
    // the two 16-byte pieces this thread gathers per chunk have a fixed (tap, ci)
    int pk_ci[2], pk_dy[2], pk_dx[2];
    bool pk_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int pc = (tid + i * WG_THREADS) & 15;
        const int n = n0 + pc * 8;
        pk_ok[i] = n < NT;
        const int tap = pk_ok[i] ? n / p.Cin : 0;
        pk_ci[i] = n - tap * p.Cin;
        const int ky = tap / p.KW, kx = tap - ky * p.KW;
        pk_dy[i] = ky * p.dil - p.pad_t;
        pk_dx[i] = kx * p.dil - p.pad_l;
    }

    auto load_chunk = [&](int chunk, int stage) {
        const uint32_t a_dst = smem_u32t(sA + stage * A_STAGE), b_dst = smem_u32t(sB + stage * B_STAGE);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int q = tid + i * WG_THREADS, k = q >> 4, pc = q & 15;
            const int m = chunk * WG_BK + k;
            const bool mv = m < p.M;
            const int mm = mv ? m : 0;
            // dY piece
            const int co = co0 + pc * 8;
            const bool va = mv && co < p.ldy;
            cp_async16t(a_dst + wg_off(k, pc), va ? p.dy + (size_t)mm * p.ldy + co : p.dy, va);
            // X piece (im2col gather for this tap)
            const int b = mm / p.HoWo, pix = mm - b * p.HoWo, oy = pix / p.Wo, ox = pix - oy * p.Wo;
            const int iy = oy * p.stride + pk_dy[i], ix = ox * p.stride + pk_dx[i];
            const int ci = pk_ci[i];
            const bool vb = mv && pk_ok[i] && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
            cp_async16t(b_dst + wg_off(k, pc), vb ? p.x + (((size_t)b * p.H + iy) * p.W + ix) * p.Cin + ci : p.x, vb);
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

    const int n_chunks = c_end - c_begin;
#pragma unroll
    for (int s = 0; s < WG_STAGES - 1; ++s) {
        if (s < n_chunks) load_chunk(c_begin + s, s);
        asm volatile("cp.async.commit_group;\n" ::);
    }
    for (int c = 0; c < n_chunks; ++c) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(WG_STAGES - 2));
        __syncthreads();
        const int nxt = c + WG_STAGES - 1;
        if (nxt < n_chunks) load_chunk(c_begin + nxt, nxt % WG_STAGES);
        asm volatile("cp.async.commit_group;\n" ::);
        const int stage = c % WG_STAGES;
        const uint32_t a_s = smem_u32t(sA + stage * A_STAGE), b_s = smem_u32t(sB + stage * B_STAGE);
#pragma unroll
        for (int ks = 0; ks < WG_BK / 16; ++ks) {
            uint32_t af[4][4], bf[2][4];
            const int mi = lane >> 3, r = lane & 7;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {               // A stored [k][m]: matrices (k0-7,m0-7)(k0-7,m8-15)(k8-15,m0-7)(k8-15,m8-15)
                const int k = ks * 16 + (mi >> 1) * 8 + r;
                const int mcol = warp_m * 64 + mt * 16 + (mi & 1) * 8;
                ldmatrix_x4_trans(af[mt], a_s + wg_off(k, mcol >> 3));
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {               // B stored [k][n]: (k0-7,n0-7)(k8-15,n0-7)(k0-7,n8-15)(k8-15,n8-15)
                const int k = ks * 16 + (mi & 1) * 8 + r;
                const int ncol = warp_n * 32 + np * 16 + (mi >> 1) * 8;
                ldmatrix_x4_trans(bf[np], b_s + wg_off(k, ncol >> 3));
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma16816t(acc[mt][nt], af[mt], bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);

    // fp32 atomics combine the K splits (and, for shared filters, several calls)
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = co0 + warp_m * 64 + mt * 16 + (lane >> 2) + h * 8;
                const int n = n0 + warp_n * 32 + nt * 8 + (lane & 3) * 2;
                if (co < p.Cout) {
                    float* dst = p.dw + (size_t)co * NT + n;
                    if (n < NT) atomicAdd(dst, acc[mt][nt][h * 2]);
                    if (n + 1 < NT) atomicAdd(dst + 1, acc[mt][nt][h * 2 + 1]);
                }
            }
}

// ---------------------------------------------------------------- helpers --
// dY *= (Y > 0): ReLU backward, in place, 8 halfs per thread.
__global__ void __launch_bounds__(256)
relu_bwd_kernel(uint4* __restrict__ dy, const uint4* __restrict__ y, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 g = dy[i];
        const uint4 v = __ldg(y + i);
        __half2* gh = reinterpret_cast<__half2*>(&g);
        const __half2* vh = reinterpret_cast<const __half2*>(&v);
        const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
        for (int k = 0; k < 4; ++k) gh[k] = __hmul2(gh[k], __hgt2(vh[k], zero));
        dy[i] = g;
    }
}

// db[c] += sum_rows dY[row][c]   (dY [rows][ld] fp16, first C columns).
// Threads map to (row slice, channel group of 8) with the channel group fastest, so a warp reads
// contiguous 16-byte pieces; slices of one block are combined in shared memory before the atomics.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __half* __restrict__ dy, float* __restrict__ db, int64_t rows, int ld, int C) {
    extern __shared__ float s_part[];                                 // [slices][ld]
    const int G = ld >> 3;                                            // channel groups per row
    const int slices = blockDim.x / G;                                // host guarantees G <= 256
    const int g = threadIdx.x % G, slice = threadIdx.x / G;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (slice < slices) {
        for (int64_t r = (int64_t)blockIdx.x * slices + slice; r < rows; r += (int64_t)gridDim.x * slices) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + r * ld) + g);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s_part[slice * ld + g * 8 + k] = acc[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.0f;
        for (int sl = 0; sl < slices; ++sl) t += s_part[sl * ld + c];
        atomicAdd(db + c, t);
    }
}

// wt[ci][KH-1-ky][KW-1-kx][co (padded to ldo)] = w[co][ky][kx][ci]
__global__ void __launch_bounds__(256)
filter_flip_transpose_kernel(const __half* __restrict__ w, __half* __restrict__ wt, int Cout, int KH, int KW, int Cin, int ldo) {
    const int64_t total = (int64_t)Cin * KH * KW * ldo;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(e % ldo);
        int64_t r = e / ldo;
        const int kx = (int)(r % KW); r /= KW;
        const int ky = (int)(r % KH);
        const int ci = (int)(r / KH);
        wt[e] = co < Cout ? w[(((size_t)co * KH + (KH - 1 - ky)) * KW + (KW - 1 - kx)) * Cin + ci] : __float2half_rn(0.0f);
    }
}

// out[b][oy*s][ox*s][:] = in[b][oy][ox][:], zeros elsewhere  (out must be pre-zeroed once; written positions are fixed)
__global__ void __launch_bounds__(256)
upsample_zero_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int Ho, int Wo, int C8, int Hu, int Wu, int s, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % C8);
        int64_t r = t / C8;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int b = (int)(r / Ho);
        out[(((size_t)b * Hu + oy * s) * Wu + ox * s) * C8 + c8] = in[t];
    }
}

// MaxPool2D backward: each input pixel gathers dY from the windows whose FIRST maximum (scan order ky, kx) it is.
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ y, const __half* __restrict__ dy,
                   __half* __restrict__ dx, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad_t, int pad_l,
                   int accumulate, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        int64_t r = t / C;
        const int ix = (int)(r % W); r /= W;
        const int iy = (int)(r % H);
        const int b = (int)(r / H);
        const float xv = __half2float(x[t]);
        float g = 0.0f;
        // windows (oy, ox) that contain (iy, ix)
        for (int oy = max(0, (iy + pad_t - k + stride) / stride); oy <= min(Ho - 1, (iy + pad_t) / stride); ++oy)
            for (int ox = max(0, (ix + pad_l - k + stride) / stride); ox <= min(Wo - 1, (ix + pad_l) / stride); ++ox) {
                const size_t o = (((size_t)b * Ho + oy) * Wo + ox) * C + c;
                if (__half2float(y[o]) != xv) continue;
                // first maximum in scan order wins: any earlier element of the window equal to the max?
                bool first = true;
                for (int ky = 0; ky < k && first; ++ky)
                    for (int kx = 0; kx < k; ++kx) {
                        const int yy = oy * stride - pad_t + ky, xx = ox * stride - pad_l + kx;
                        if (yy == iy && xx == ix) { ky = k; break; }
                        if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W &&
                            __half2float(x[(((size_t)b * H + yy) * W + xx) * C + c]) == xv) { first = false; break; }
                    }
                if (first) g += __half2float(dy[o]);
            }
        dx[t] = __float2half_rn(accumulate ? g + __half2float(dx[t]) : g);
    }
}

// Non-overlapping windows (k == stride, the four 2x2/2 pools): one thread per output window and 8
// channels; the first maximum of the window (scan order) receives dY, every other input of the
// window receives 0, so each dX element is written exactly once with 16-byte accesses.
template <int K>
__global__ void __launch_bounds__(256)
maxpool_bwd_disjoint_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                            int H, int W, int C8, int Ho, int Wo, int pad_t, int pad_l, int accumulate, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % C8);
        int64_t r = t / C8;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const uint4 gv = __ldg(dy + t);
        const __half* gh = reinterpret_cast<const __half*>(&gv);
        uint4 xv[K * K];
        bool ok[K * K];
        float best[8];
        int arg[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = -1; }
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int iy = oy * K - pad_t + ky, ix = ox * K - pad_l + kx;
                const int q = ky * K + kx;
                ok[q] = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
                if (ok[q]) {
                    xv[q] = __ldg(x + (((size_t)b * H + iy) * W + ix) * C8 + c8);
                    const __half* xh = reinterpret_cast<const __half*>(&xv[q]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float v = __half2float(xh[e]);
                        if (v > best[e]) { best[e] = v; arg[e] = q; }        // strict: the first maximum wins
                    }
                }
            }
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int q = ky * K + kx;
                if (!ok[q]) continue;
                const int iy = oy * K - pad_t + ky, ix = ox * K - pad_l + kx;
                uint4* dst = dx + (((size_t)b * H + iy) * W + ix) * C8 + c8;
                uint4 o;
                __half* oh = reinterpret_cast<__half*>(&o);
                uint4 prev = make_uint4(0, 0, 0, 0);
                if (accumulate) prev = *dst;
                const __half* ph = reinterpret_cast<const __half*>(&prev);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float g = arg[e] == q ? __half2float(gh[e]) : 0.0f;
                    oh[e] = __float2half_rn(accumulate ? g + __half2float(ph[e]) : g);
                }
                *dst = o;
            }
    }
}

// L2Normalization backward (models/ssd_vgg16.py:63): y = x * r * s,  r = rsqrt(max(sum x^2, eps)).
// dx = s*r*dy - x * r^3 * sum_c(dy*s*x)   (when sum x^2 > eps);  dscale[c] += sum_rows dy * x * r.
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const __half* __restrict__ x, const float* __restrict__ scale, const __half* __restrict__ dy,
                  __half* __restrict__ dx, float* __restrict__ dscale, int64_t rows, int C, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const __half* xr = x + r * C;
        const __half* gr = dy + r * C;
        float ss = 0.0f, dot = 0.0f;
        for (int c = lane; c < C; c += 32) {
            const float xv = __half2float(xr[c]), gv = __half2float(gr[c]);
            ss = fmaf(xv, xv, ss);
            dot = fmaf(gv * __ldg(scale + c), xv, dot);
        }
        ss = warp_sum(ss); dot = warp_sum(dot);
        const bool clamped = ss < 1e-12f;
        const float inv = rsqrtf(fmaxf(ss, 1e-12f));
        const float coef = clamped ? 0.0f : dot * inv * inv * inv;
        for (int c = lane; c < C; c += 32) {
            const float xv = __half2float(xr[c]), gv = __half2float(gr[c]);
            float g = gv * __ldg(scale + c) * inv - xv * coef;
            if (accumulate) g += __half2float(dx[r * C + c]);
            dx[r * C + c] = __float2half_rn(g);
            atomicAdd(dscale + c, gv * xv * inv);
        }
    }
}

// Head gradient gather: dY[b][pix][0..A*L) <- g_logits, [A*L .. A*(L+4)) <- g_deltas, zero padding to ld.
__global__ void __launch_bounds__(256)
head_grad_gather_kernel(const float* __restrict__ g_logits, const float* __restrict__ g_deltas, __half* __restrict__ dy,
                        int N, int L, int off, int HW, int A, int ld, int64_t total) {
    const int AL = A * L, A4 = A * 4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % ld);
        const int64_t r = t / ld;
        const int pix = (int)(r % HW), b = (int)(r / HW);
        float v = 0.0f;
        if (c < AL) v = g_logits[((size_t)b * N + off) * L + (size_t)pix * AL + c];
        else if (c < AL + A4) v = g_deltas[((size_t)b * N + off) * 4 + (size_t)pix * A4 + (c - AL)];
        dy[t] = __float2half_rn(v);
    }
}

// Fused Adam (Keras defaults, trainer.py:92): g = grad * inv_scale + l2 * w;  m, v update;  w -= lr_t * m / (sqrt(v) + eps).
// Also refreshes the fp16 working copy and accumulates sum(w^2) of the pre-update weights (regulariser loss).
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
            __half* __restrict__ w16, int64_t n, float lr_t, float b1, float b2, float eps, float inv_scale, float l2,
            float* __restrict__ sumsq) {
    float local = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float wv = w[i];
        const float gv = g[i] * inv_scale + l2 * wv;
        const float mv = b1 * m[i] + (1.0f - b1) * gv;
        const float vv = b2 * v[i] + (1.0f - b2) * gv * gv;
        const float nw = wv - lr_t * mv / (sqrtf(vv) + eps);
        m[i] = mv; v[i] = vv; w[i] = nw;
        if (w16) w16[i] = __float2half_rn(nw);
        local += wv * wv;
    }
    if (sumsq) {
        local = warp_sum(local);
        if ((threadIdx.x & 31) == 0 && local != 0.0f) atomicAdd(sumsq, local);
    }
}

// Multi-tensor variant: one launch updates every variable of the model.  blockIdx.y selects the descriptor,
// blockIdx.x strides over its elements (descriptors shorter than the grid simply finish early).
// ``guard`` (may be NULL): guard[0] != 0 means this step's gradients hold a non-finite value (fp16 overflow of a
// loss-scaled activation gradient): the whole update is skipped -- weights, moments and the fp16 copies stay as they
// are -- and guard[1] counts the skipped steps for the host's loss-scale controller.
__global__ void __launch_bounds__(256)
adam_multi_kernel(const ssd_adam_var* __restrict__ vars, float lr_t, float b1, float b2, float eps, float inv_scale,
                  float* __restrict__ sumsq, int* __restrict__ guard) {
    if (guard && guard[0] != 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(guard + 1, 1);
        return;
    }
    const ssd_adam_var d = vars[blockIdx.y];
    float* __restrict__ w = d.w; float* __restrict__ m = d.m; float* __restrict__ v = d.v;
    const float* __restrict__ g = d.grad;
    __half* __restrict__ w16 = reinterpret_cast<__half*>(d.w16);
    float local = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float wv = w[i];
        const float gv = g[i] * inv_scale + d.l2 * wv;
        const float mv = b1 * m[i] + (1.0f - b1) * gv;
        const float vv = b2 * v[i] + (1.0f - b2) * gv * gv;
        const float nw = wv - lr_t * mv / (sqrtf(vv) + eps);
        m[i] = mv; v[i] = vv; w[i] = nw;
        if (w16) w16[i] = __float2half_rn(nw);
        local += wv * wv;
    }
    if (sumsq && d.l2 != 0.0f) {
        local = warp_sum(local);
        if ((threadIdx.x & 31) == 0 && local != 0.0f) atomicAdd(sumsq, local);
    }
}

// Any non-finite gradient among all variables -> flag[0] = 1 (the caller zeroes it first).
__global__ void __launch_bounds__(256)
grad_nonfinite_kernel(const ssd_adam_var* __restrict__ vars, int* __restrict__ flag) {
    const ssd_adam_var d = vars[blockIdx.y];
    const float* __restrict__ g = d.grad;
    bool bad = false;
    const int64_t n4 = d.n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);          // gradient views start on 16-byte boundaries
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        // x - x is 0 for finite x and NaN for +-inf / NaN
        const float t = (v.x - v.x) + (v.y - v.y) + (v.z - v.z) + (v.w - v.w);
        bad |= !(t == 0.0f);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(d.n & 3)) {
        const float x = g[(n4 << 2) + threadIdx.x];
        bad |= !((x - x) == 0.0f);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// conv_tcgen05.cu
bool conv_wgrad_tcgen05_supported(const ssd_conv_desc* d, int ldy);
int  conv_wgrad_tcgen05_launch(const ssd_conv_desc* d, const void* d_dy, int ldy, float* d_dw, cudaStream_t st);

static bool force_legacy_wgrad() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SSD_B200_CONV");
        v = (e && strcmp(e, "legacy") == 0) ? 1 : 0;
    }
    return v == 1;
}

static int grid1d(int64_t threads, int per_sm = 8) {
    int64_t blocks = (threads + 255) / 256, cap = (int64_t)sm_count() * per_sm;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_conv2d_wgrad(const ssd_conv_desc* d, const void* d_dy, int ldy, float* d_dw, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d); SSD_REQUIRE_PTR(d->in); SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_dw);
    SSD_REQUIRE(d->B >= 1 && d->H >= 1 && d->W >= 1 && d->Cin >= 8 && d->Cin % 8 == 0 && d->Cout >= 1 && d->Ho >= 1 &&
                d->Wo >= 1 && d->KH >= 1 && d->KW >= 1 && d->stride >= 1 && d->dilation >= 1 && ldy >= d->Cout && ldy % 8 == 0,
                SSD_ERR_SHAPE, "ssd_conv2d_wgrad: bad shape (Cin=%d Cout=%d ldy=%d; Cin and ldy must be multiples of 8)",
                d->Cin, d->Cout, ldy);
    if (!force_legacy_wgrad() && conv_wgrad_tcgen05_supported(d, ldy))
        return conv_wgrad_tcgen05_launch(d, d_dy, ldy, d_dw, as_stream(stream));
    WgradK k;
    k.x = (const __half*)d->in; k.dy = (const __half*)d_dy; k.dw = d_dw;
    k.B = d->B; k.H = d->H; k.W = d->W; k.Cin = d->Cin; k.Ho = d->Ho; k.Wo = d->Wo; k.Cout = d->Cout; k.KW = d->KW;
    k.stride = d->stride; k.dil = d->dilation; k.pad_t = d->pad_top; k.pad_l = d->pad_left;
    k.HoWo = d->Ho * d->Wo; k.M = d->B * k.HoWo; k.ldy = ldy; k.taps = d->KH * d->KW;
    k.chunks_total = (k.M + WG_BK - 1) / WG_BK;
    const int tiles_m = (d->Cout + WG_BM - 1) / WG_BM;
    k.tiles_n = (k.taps * d->Cin + WG_BN - 1) / WG_BN;
    const int base = tiles_m * k.tiles_n;
    int splits = (sm_count() * 2 + base - 1) / base;                       // ~2 CTAs per SM overall
    splits = max(1, min(splits, (k.chunks_total + 7) / 8));
    k.chunks_per_split = (k.chunks_total + splits - 1) / splits;
    splits = (k.chunks_total + k.chunks_per_split - 1) / k.chunks_per_split;
    const size_t smem = (size_t)WG_STAGES * WG_BK * (WG_BM + WG_BN) * 2;
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "ssd_conv2d_wgrad: cudaFuncSetAttribute");
    dim3 grid(tiles_m * k.tiles_n, 1, splits);
    conv_wgrad_kernel<<<grid, WG_THREADS, smem, as_stream(stream)>>>(k);
    SSD_CHECK_LAUNCH("conv_wgrad_kernel");
    return SSD_OK;
}

extern "C" int ssd_relu_bwd(void* d_dy, const void* d_y, int64_t n, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_y);
    SSD_REQUIRE(n >= 0 && n % 8 == 0, SSD_ERR_SHAPE, "ssd_relu_bwd: n=%lld must be a multiple of 8", (long long)n);
    if (n == 0) return SSD_OK;
    relu_bwd_kernel<<<grid1d(n / 8, 16), 256, 0, as_stream(stream)>>>(reinterpret_cast<uint4*>(d_dy),
                                                                      reinterpret_cast<const uint4*>(d_y), n / 8);
    SSD_CHECK_LAUNCH("relu_bwd_kernel");
    return SSD_OK;
}

extern "C" int ssd_bias_grad(const void* d_dy, float* d_db, int64_t rows, int ld, int C, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_db);
    SSD_REQUIRE(rows >= 0 && ld % 8 == 0 && C >= 1 && C <= ld, SSD_ERR_SHAPE, "ssd_bias_grad: bad shape rows=%lld ld=%d C=%d",
                (long long)rows, ld, C);
    if (rows == 0) return SSD_OK;
    SSD_REQUIRE(ld <= 2048, SSD_ERR_UNSUPPORTED, "ssd_bias_grad: ld=%d > 2048", ld);
    const int G = ld / 8, slices = 256 / G;
    const int64_t want = (rows + slices - 1) / slices, cap = (int64_t)sm_count() * 8;
    const size_t smem = (size_t)slices * ld * sizeof(float);
    bias_grad_kernel<<<(int)(want < cap ? want : cap), 256, smem, as_stream(stream)>>>((const __half*)d_dy, d_db, rows, ld, C);
    SSD_CHECK_LAUNCH("bias_grad_kernel");
    return SSD_OK;
}

extern "C" int ssd_filter_flip_transpose(const void* d_w, void* d_wt, int Cout, int KH, int KW, int Cin, int ldo,
                                         ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_w); SSD_REQUIRE_PTR(d_wt);
    SSD_REQUIRE(Cout >= 1 && KH >= 1 && KW >= 1 && Cin >= 1 && ldo >= Cout, SSD_ERR_SHAPE, "ssd_filter_flip_transpose: bad shape");
    filter_flip_transpose_kernel<<<grid1d((int64_t)Cin * KH * KW * ldo, 8), 256, 0, as_stream(stream)>>>(
        (const __half*)d_w, (__half*)d_wt, Cout, KH, KW, Cin, ldo);
    SSD_CHECK_LAUNCH("filter_flip_transpose_kernel");
    return SSD_OK;
}

extern "C" int ssd_upsample_zero(const void* d_in, void* d_out, int B, int Ho, int Wo, int C, int Hu, int Wu, int s,
                                 ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_in); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(B >= 1 && Ho >= 1 && Wo >= 1 && C % 8 == 0 && s >= 1 && Hu >= (Ho - 1) * s + 1 && Wu >= (Wo - 1) * s + 1,
                SSD_ERR_SHAPE, "ssd_upsample_zero: bad shape");
    const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
    upsample_zero_kernel<<<grid1d(total, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_out), Ho, Wo, C / 8, Hu, Wu, s, total);
    SSD_CHECK_LAUNCH("upsample_zero_kernel");
    return SSD_OK;
}

extern "C" int ssd_maxpool_bwd(const void* d_x, const void* d_y, const void* d_dy, void* d_dx, int B, int H, int W, int C,
                               int Ho, int Wo, int k, int stride, int pad_top, int pad_left, int accumulate,
                               ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_x); SSD_REQUIRE_PTR(d_y); SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_dx);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1 && k >= 1 && k <= 7 && stride >= 1, SSD_ERR_SHAPE, "ssd_maxpool_bwd: bad shape");
    // disjoint windows that cover every input pixel: vectorised single-pass kernel
    if (k == stride && k == 2 && C % 8 == 0 && Ho * k - pad_top >= H && Wo * k - pad_left >= W && pad_top >= 0 && pad_left >= 0) {
        const int64_t total8 = (int64_t)B * Ho * Wo * (C / 8);
        maxpool_bwd_disjoint_kernel<2><<<grid1d(total8, 16), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const uint4*>(d_x), reinterpret_cast<const uint4*>(d_dy), reinterpret_cast<uint4*>(d_dx),
            H, W, C / 8, Ho, Wo, pad_top, pad_left, accumulate, total8);
        SSD_CHECK_LAUNCH("maxpool_bwd_disjoint_kernel");
        return SSD_OK;
    }
    const int64_t total = (int64_t)B * H * W * C;
    maxpool_bwd_kernel<<<grid1d(total, 16), 256, 0, as_stream(stream)>>>(
        (const __half*)d_x, (const __half*)d_y, (const __half*)d_dy, (__half*)d_dx, H, W, C, Ho, Wo, k, stride, pad_top,
        pad_left, accumulate, total);
    SSD_CHECK_LAUNCH("maxpool_bwd_kernel");
    return SSD_OK;
}

extern "C" int ssd_l2norm_bwd(const void* d_x, const float* d_scale, const void* d_dy, void* d_dx, float* d_dscale,
                              int64_t rows, int C, int accumulate, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_x); SSD_REQUIRE_PTR(d_scale); SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_dx); SSD_REQUIRE_PTR(d_dscale);
    SSD_REQUIRE(rows >= 0 && C >= 1, SSD_ERR_SHAPE, "ssd_l2norm_bwd: bad shape");
    if (rows == 0) return SSD_OK;
    l2norm_bwd_kernel<<<grid1d(rows * 32, 8), 256, 0, as_stream(stream)>>>(
        (const __half*)d_x, d_scale, (const __half*)d_dy, (__half*)d_dx, d_dscale, rows, C, accumulate);
    SSD_CHECK_LAUNCH("l2norm_bwd_kernel");
    return SSD_OK;
}

extern "C" int ssd_head_grad_gather(const float* d_g_logits, const float* d_g_deltas, void* d_dy, int B, int N, int L,
                                    int anchor_offset, int HW, int A, int ld, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_g_logits); SSD_REQUIRE_PTR(d_g_deltas); SSD_REQUIRE_PTR(d_dy);
    SSD_REQUIRE(B >= 1 && N >= 1 && L >= 1 && HW >= 1 && A >= 1 && ld >= A * (L + 4) && anchor_offset >= 0 &&
                anchor_offset + HW * A <= N, SSD_ERR_SHAPE, "ssd_head_grad_gather: bad shape");
    const int64_t total = (int64_t)B * HW * ld;
    head_grad_gather_kernel<<<grid1d(total, 8), 256, 0, as_stream(stream)>>>(d_g_logits, d_g_deltas, (__half*)d_dy, N, L,
                                                                             anchor_offset, HW, A, ld, total);
    SSD_CHECK_LAUNCH("head_grad_gather_kernel");
    return SSD_OK;
}

extern "C" int ssd_adam_step(float* d_w, float* d_m, float* d_v, const float* d_grad, void* d_w16, int64_t n, float lr_t,
                             float beta1, float beta2, float eps, float inv_scale, float l2, float* d_sumsq,
                             ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_w); SSD_REQUIRE_PTR(d_m); SSD_REQUIRE_PTR(d_v); SSD_REQUIRE_PTR(d_grad);
    SSD_REQUIRE(n >= 0, SSD_ERR_SHAPE, "ssd_adam_step: n=%lld", (long long)n);
    if (n == 0) return SSD_OK;
    adam_kernel<<<grid1d(n, 8), 256, 0, as_stream(stream)>>>(d_w, d_m, d_v, d_grad, (__half*)d_w16, n, lr_t, beta1, beta2, eps,
                                                            inv_scale, l2, d_sumsq);
    SSD_CHECK_LAUNCH("adam_kernel");
    return SSD_OK;
}

static int adam_multi_grid(int n_vars, int64_t max_n) {
    // enough CTAs per variable that the largest one (a few million elements) still spreads over the device
    int64_t bx = (max_n + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = max((int64_t)1, (int64_t)sm_count() * 16 / n_vars + 1);
    if (bx > cap) bx = cap;
    return (int)(bx < 1 ? 1 : bx);
}

extern "C" int ssd_grad_nonfinite_multi(const ssd_adam_var* d_vars, int n_vars, int64_t max_n, int* d_flag, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_vars); SSD_REQUIRE_PTR(d_flag);
    SSD_REQUIRE(n_vars >= 0 && n_vars <= 65535 && max_n >= 0, SSD_ERR_SHAPE, "ssd_grad_nonfinite_multi: n_vars=%d max_n=%lld",
                n_vars, (long long)max_n);
    cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(int), as_stream(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ssd_grad_nonfinite_multi: memset");
    if (n_vars == 0 || max_n == 0) return SSD_OK;
    grad_nonfinite_kernel<<<dim3((unsigned)adam_multi_grid(n_vars, max_n), (unsigned)n_vars), 256, 0, as_stream(stream)>>>(d_vars, d_flag);
    SSD_CHECK_LAUNCH("grad_nonfinite_kernel");
    return SSD_OK;
}

extern "C" int ssd_adam_step_multi_guarded(const ssd_adam_var* d_vars, int n_vars, int64_t max_n, float lr_t, float beta1,
                                           float beta2, float eps, float inv_scale, float* d_sumsq, int* d_guard,
                                           ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_vars);
    SSD_REQUIRE(n_vars >= 0 && n_vars <= 65535 && max_n >= 0, SSD_ERR_SHAPE, "ssd_adam_step_multi_guarded: n_vars=%d max_n=%lld",
                n_vars, (long long)max_n);
    if (n_vars == 0 || max_n == 0) return SSD_OK;
    adam_multi_kernel<<<dim3((unsigned)adam_multi_grid(n_vars, max_n), (unsigned)n_vars), 256, 0, as_stream(stream)>>>(
        d_vars, lr_t, beta1, beta2, eps, inv_scale, d_sumsq, d_guard);
    SSD_CHECK_LAUNCH("adam_multi_kernel");
    return SSD_OK;
}

extern "C" int ssd_adam_step_multi(const ssd_adam_var* d_vars, int n_vars, int64_t max_n, float lr_t, float beta1, float beta2,
                                   float eps, float inv_scale, float* d_sumsq, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_vars);
    SSD_REQUIRE(n_vars >= 0 && n_vars <= 65535 && max_n >= 0, SSD_ERR_SHAPE, "ssd_adam_step_multi: n_vars=%d max_n=%lld", n_vars,
                (long long)max_n);
    if (n_vars == 0 || max_n == 0) return SSD_OK;
    // enough CTAs per variable that the largest one (a few million elements) still spreads over the device
    int64_t bx = (max_n + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = max((int64_t)1, (int64_t)sm_count() * 16 / n_vars + 1);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)n_vars), 256, 0, as_stream(stream)>>>(d_vars, lr_t, beta1, beta2, eps,
                                                                                          inv_scale, d_sumsq, nullptr);
    SSD_CHECK_LAUNCH("adam_multi_kernel");
    return SSD_OK;
}
