// Training-mode BatchNormalization and the depthwise-convolution gradients: what the
// MobileNetV2 backbone (models/ssd_mobilenet_v2.py:25 -> keras_applications MobileNetV2:
// Conv2D/DepthwiseConv2D without bias -> BatchNormalization(epsilon=1e-3, momentum=0.999) -> ReLU6)
// needs on top of train_kernels.cu for the Keras fit step of trainer.py:86-127.
//
// All of these are HBM-bound streaming kernels over NHWC fp16 activations viewed as an
// [M = B*H*W, C] matrix of 16-byte channel groups (C % 8 == 0):
//
//   * per-channel reductions (batch statistics, dgamma/dbeta, depthwise filter gradient) use a
//     flat decomposition in which a thread's channel group never changes (the thread stride is a
//     multiple of C/8), accumulate in fp32 registers, combine inside the CTA through shared memory
//     in a fixed order and leave one partial row per CTA; a tiny second kernel adds the partial rows
//     in order (double precision) -- deterministic, no atomics on the statistics;
//   * element-wise passes (normalise + activation + residual, dX of BatchNorm, depthwise dX) are
//     16-byte vector loads/stores with fp32 math.

#include "common.cuh"

namespace ssd {

constexpr int kRedThreads = 256;
constexpr int kMaxChunks = 592;                 // 148 SMs x 4 resident CTAs

__device__ __forceinline__ void h8_unpack(const uint4 v, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 h8_pack(const float (&f)[8]) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

struct ChunkGeom {
    int C8;                 // channel groups per row
    int tpr;                // active threads per CTA: largest multiple of C8 <= kRedThreads
    int64_t total;          // M * C8 vector elements
    int64_t per_chunk;      // vector elements per CTA (multiple of tpr)
    int chunks;
};

static ChunkGeom chunk_geom(int64_t M, int C, int per_thread = 16) {
    ChunkGeom g;
    g.C8 = C / 8;
    g.tpr = (kRedThreads / g.C8) * g.C8;
    g.total = M * g.C8;
    int64_t want = (g.total + (int64_t)g.tpr * per_thread - 1) / ((int64_t)g.tpr * per_thread);   // >= per_thread elements each
    g.chunks = (int)(want < 1 ? 1 : want > kMaxChunks ? kMaxChunks : want);
    int64_t per = (g.total + g.chunks - 1) / g.chunks;
    g.per_chunk = (per + g.tpr - 1) / g.tpr * g.tpr;
    g.chunks = (int)((g.total + g.per_chunk - 1) / g.per_chunk);
    return g;
}

// CTA-level fixed-order combine of NACC fp32 accumulators per thread: threads with the same channel
// group (t % C8) are added in increasing t.  Result rows go to out[(slot * C8 + cg) * 8 + k].
template <int NSLOT>
__device__ __forceinline__ void cta_combine(float (&acc)[NSLOT][8], int C8, int tpr, float* s_red, float* out_row) {
    // s_red: [kRedThreads][NSLOT*8]
    const int t = threadIdx.x;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) s_red[(size_t)t * (NSLOT * 8) + s * 8 + k] = acc[s][k];
    __syncthreads();
    for (int o = t; o < C8 * NSLOT * 8; o += kRedThreads) {
        const int cg = o / (NSLOT * 8), r = o - cg * (NSLOT * 8);
        float v = 0.0f;
        for (int j = cg; j < tpr; j += C8) v += s_red[(size_t)j * (NSLOT * 8) + r];
        const int s = r >> 3, k = r & 7;
        out_row[((size_t)s * C8 + cg) * 8 + k] = v;
    }
}

// ------------------------------------------------------------- BN statistics --
__global__ void __launch_bounds__(kRedThreads)
bn_stats_partial_kernel(const uint4* __restrict__ x, ChunkGeom g, float* __restrict__ partial) {
    extern __shared__ float s_red[];
    float acc[2][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[0][k] = 0.f; acc[1][k] = 0.f; }
    const int64_t lo = (int64_t)blockIdx.x * g.per_chunk;
    const int64_t hi = lo + g.per_chunk < g.total ? lo + g.per_chunk : g.total;
    if ((int)threadIdx.x < g.tpr) {
        for (int64_t e = lo + threadIdx.x; e < hi; e += g.tpr) {
            float f[8];
            h8_unpack(__ldg(x + e), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc[0][k] += f[k]; acc[1][k] = fmaf(f[k], f[k], acc[1][k]); }
        }
    }
    cta_combine<2>(acc, g.C8, g.tpr, s_red, partial + (size_t)blockIdx.x * 2 * g.C8 * 8);
}

// Fixed-order sum of column c of the partial rows by ONE WARP: lane l adds rows l, l+32, ... (double), then a
// shuffle tree -- the same association every run, and 32 independent load streams instead of one serial chain.
__device__ __forceinline__ double warp_column_sum(const float* __restrict__ col, int chunks, size_t row_stride, int lane) {
    double s = 0.0;
    for (int j = lane; j < chunks; j += 32) s += (double)__ldg(col + (size_t)j * row_stride);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// One warp per channel: mean / biased variance -> save[0..C) = mean, save[C..2C) = rstd; moving statistics
// updated like Keras' fused BatchNormalization ([TF-recall]: moving_mean = m*mom + mean*(1-mom);
// moving_variance uses the unbiased batch variance).
__global__ void __launch_bounds__(256)
bn_stats_finalize_kernel(const float* __restrict__ partial, int chunks, int C, double inv_m, double bessel,
                         float eps, float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ save, float* __restrict__ coef,
                         float* __restrict__ moving_mean, float* __restrict__ moving_var) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    const double s = warp_column_sum(partial + c, chunks, (size_t)2 * C, lane);
    const double ss = warp_column_sum(partial + C + c, chunks, (size_t)2 * C, lane);
    if (lane) return;
    const double mean = s * inv_m;
    double var = ss * inv_m - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    save[c] = (float)mean;
    save[C + c] = rstd;
    const float a = gamma[c] * rstd;                  // y = a * x + b
    coef[c] = a;
    coef[C + c] = fmaf(-(float)mean, a, beta[c]);
    if (moving_mean) moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.0f - momentum);
    if (moving_var) moving_var[c] = moving_var[c] * momentum + (float)(var * bessel) * (1.0f - momentum);
}

// y = act(a * x + b) (+ res) with the per-channel affine of the finalize kernel.  The thread stride is a
// multiple of C8, so a thread's channel group -- and its 16 coefficients -- never change.
__global__ void __launch_bounds__(256)
bn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ coef, const uint4* __restrict__ res,
                uint4* __restrict__ y, int C8, int C, int act, int64_t total, int64_t stride) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    const int cg = (int)(t % C8);
    float a[8], b[8];
    {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(coef) + cg * 2), a1 = __ldg(reinterpret_cast<const float4*>(coef) + cg * 2 + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(coef + C) + cg * 2), b1 = __ldg(reinterpret_cast<const float4*>(coef + C) + cg * 2 + 1);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    }
    const float lo = act == SSD_ACT_NONE ? -__int_as_float(0x7f800000) : 0.0f;
    const float hi = act == SSD_ACT_RELU6 ? 6.0f : __int_as_float(0x7f800000);
    for (int64_t e = t; e < total; e += stride) {
        float f[8];
        h8_unpack(__ldcs(x + e), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fminf(fmaxf(fmaf(f[k], a[k], b[k]), lo), hi);
        if (res) {
            float r[8];
            h8_unpack(__ldg(res + e), r);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += r[k];
        }
        y[e] = h8_pack(f);
    }
}

// --------------------------------------------------------------- BN backward --
// g = dy * act'(bn(x));  partial sums of g and g * xhat per channel (thread-constant channel group).
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_partial_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ save, int C, int act, ChunkGeom g,
                      float* __restrict__ partial) {
    extern __shared__ float s_red[];
    float acc[2][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[0][k] = 0.f; acc[1][k] = 0.f; }
    const int64_t lo = (int64_t)blockIdx.x * g.per_chunk;
    const int64_t hi = lo + g.per_chunk < g.total ? lo + g.per_chunk : g.total;
    if ((int)threadIdx.x < g.tpr) {
        const int cg = (int)((lo + threadIdx.x) % g.C8);
        float mean[8], rstd[8], gam[8], bet[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = cg * 8 + k;
            mean[k] = __ldg(save + c); rstd[k] = __ldg(save + C + c); gam[k] = __ldg(gamma + c); bet[k] = __ldg(beta + c);
        }
        for (int64_t e = lo + threadIdx.x; e < hi; e += g.tpr) {
            float xf[8], df[8];
            h8_unpack(__ldg(x + e), xf);
            h8_unpack(__ldg(dy + e), df);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = (xf[k] - mean[k]) * rstd[k];
                float gv = df[k];
                if (act != SSD_ACT_NONE) {
                    const float v = fmaf(xh, gam[k], bet[k]);
                    if (!(v > 0.0f) || (act == SSD_ACT_RELU6 && !(v < 6.0f))) gv = 0.0f;
                }
                acc[0][k] += gv;
                acc[1][k] = fmaf(gv, xh, acc[1][k]);
            }
        }
    }
    cta_combine<2>(acc, g.C8, g.tpr, s_red, partial + (size_t)blockIdx.x * 2 * g.C8 * 8);
}

// dbeta += sum g, dgamma += sum g*xhat; coef[0..C) = sum g / M, coef[C..2C) = sum g*xhat / M  (one warp per channel)
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int chunks, int C, double inv_m,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    const double s1 = warp_column_sum(partial + c, chunks, (size_t)2 * C, lane);
    const double s2 = warp_column_sum(partial + C + c, chunks, (size_t)2 * C, lane);
    if (lane) return;
    if (dbeta) dbeta[c] += (float)s1;
    if (dgamma) dgamma[c] += (float)s2;
    coef[c] = (float)(s1 * inv_m);
    coef[C + c] = (float)(s2 * inv_m);
}

// dx = gamma * rstd * (g - mean(g) - xhat * mean(g*xhat));  dres (+)= dy.  Thread-constant channel group.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ save, const float* __restrict__ coef,
                    int C8, int C, int act, uint4* __restrict__ dx, uint4* __restrict__ dres, int accumulate_res,
                    int64_t total, int64_t stride) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    const int cg = (int)(t % C8);
    float mean[8], rstd[8], gam[8], bet[8], k1[8], k2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cg * 8 + k;
        mean[k] = __ldg(save + c); rstd[k] = __ldg(save + C + c); gam[k] = __ldg(gamma + c); bet[k] = __ldg(beta + c);
        k1[k] = __ldg(coef + c); k2[k] = __ldg(coef + C + c);
    }
    for (int64_t e = t; e < total; e += stride) {
        float xf[8], df[8], out[8];
        h8_unpack(__ldcs(x + e), xf);
        const uint4 dyv = __ldcs(dy + e);
        h8_unpack(dyv, df);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float xh = (xf[k] - mean[k]) * rstd[k];
            float g = df[k];
            if (act != SSD_ACT_NONE) {
                const float v = fmaf(xh, gam[k], bet[k]);
                if (!(v > 0.0f) || (act == SSD_ACT_RELU6 && !(v < 6.0f))) g = 0.0f;
            }
            out[k] = gam[k] * rstd[k] * (g - k1[k] - xh * k2[k]);
        }
        dx[e] = h8_pack(out);
        if (dres) {
            if (accumulate_res) {
                float rf[8];
                h8_unpack(dres[e], rf);
#pragma unroll
                for (int k = 0; k < 8; ++k) rf[k] += df[k];
                dres[e] = h8_pack(rf);
            } else {
                dres[e] = dyv;
            }
        }
    }
}

// ------------------------------------------------------- depthwise gradients --
// dX[b,iy,ix,c] = sum_{ky,kx} dY[b,oy,ox,c] * w[ky,kx,c]   with  oy*s - pad_t + ky == iy
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ w, uint4* __restrict__ dx, int H, int W, int C8,
                int Ho, int Wo, int stride, int pad_t, int pad_l, int accumulate, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int cg = (int)(t % C8);
        int64_t r = t / C8;
        const int ix = (int)(r % W); r /= W;
        const int iy = (int)(r % H);
        const int b = (int)(r / H);
        float acc[8];
        if (accumulate) h8_unpack(dx[t], acc);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = iy + pad_t - ky;
            if (ty < 0 || ty % stride) continue;
            const int oy = ty / stride;
            if (oy >= Ho) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = ix + pad_l - kx;
                if (tx < 0 || tx % stride) continue;
                const int ox = tx / stride;
                if (ox >= Wo) continue;
                float g[8], wf[8];
                h8_unpack(__ldg(dy + ((size_t)(b * Ho + oy) * Wo + ox) * C8 + cg), g);
                h8_unpack(__ldg(w + (size_t)(ky * 3 + kx) * C8 + cg), wf);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(g[k], wf[k], acc[k]);
            }
        }
        dx[t] = h8_pack(acc);
    }
}

// dW[ky,kx,c] += sum_{b,oy,ox} dY[b,oy,ox,c] * X[b, oy*s-pad_t+ky, ox*s-pad_l+kx, c]
// The 9 taps are processed in three launches-worth of registers (3 taps of one filter row at a time)
// to keep the accumulator count at 24 per thread.
__global__ void __launch_bounds__(kRedThreads)
dw_wgrad_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int H, int W, int Ho, int Wo, int stride,
                int pad_t, int pad_l, ChunkGeom g, float* __restrict__ dw) {
    extern __shared__ float s_red[];
    const int64_t lo = (int64_t)blockIdx.x * g.per_chunk;
    const int64_t hi = lo + g.per_chunk < g.total ? lo + g.per_chunk : g.total;
    const int C = g.C8 * 8;
    for (int ky = 0; ky < 3; ++ky) {
        float acc[3][8];
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[s][k] = 0.0f;
        if ((int)threadIdx.x < g.tpr) {
            for (int64_t e = lo + threadIdx.x; e < hi; e += g.tpr) {
                const int cg = (int)(e % g.C8);
                int64_t r = e / g.C8;
                const int ox = (int)(r % Wo); r /= Wo;
                const int oy = (int)(r % Ho);
                const int b = (int)(r / Ho);
                const int iy = oy * stride - pad_t + ky;
                if ((unsigned)iy >= (unsigned)H) continue;
                float gf[8];
                h8_unpack(__ldg(dy + e), gf);
                const uint4* row = x + ((size_t)(b * H + iy) * W) * g.C8 + cg;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = ox * stride - pad_l + kx;
                    if ((unsigned)ix >= (unsigned)W) continue;
                    float xf[8];
                    h8_unpack(__ldg(row + (size_t)ix * g.C8), xf);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[kx][k] = fmaf(gf[k], xf[k], acc[kx][k]);
                }
            }
        }
        // combine inside the CTA in a fixed order, then one atomic per (tap, channel) and CTA
        const int t = threadIdx.x;
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) s_red[(size_t)t * 24 + s * 8 + k] = acc[s][k];
        __syncthreads();
        for (int o = t; o < g.C8 * 24; o += kRedThreads) {
            const int cg = o / 24, r = o - cg * 24;
            float v = 0.0f;
            for (int j = cg; j < g.tpr; j += g.C8) v += s_red[(size_t)j * 24 + r];
            const int kx = r >> 3, k = r & 7;
            atomicAdd(dw + (size_t)(ky * 3 + kx) * C + cg * 8 + k, v);
        }
        __syncthreads();
    }
}

static int ew_grid(int64_t total) {
    int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
    return (int)(blocks < cap ? blocks : cap);
}
// grid and thread stride (a multiple of C8, at most grid*256) for the thread-constant-channel element-wise passes
static int ew_grid_c8(int64_t total, int C8, int64_t* stride) {
    int64_t blocks = (total + 4 * 256 - 1) / (4 * 256), cap = (int64_t)sm_count() * 8;     // ~4+ vectors per thread
    blocks = blocks < 1 ? 1 : blocks > cap ? cap : blocks;
    int64_t threads = blocks * 256;
    if (threads < C8) { blocks = (C8 + 255) / 256; threads = blocks * 256; }
    *stride = threads / C8 * C8;
    return (int)blocks;
}

}  // namespace ssd

using namespace ssd;

extern "C" size_t ssd_bn_workspace_bytes(int C) {
    if (C < 8) C = 8;
    // partial rows [kMaxChunks][2][C] + coefficients [2][C]
    return ((size_t)kMaxChunks * 2 * C + 2 * (size_t)C) * sizeof(float);
}

extern "C" int ssd_bn_train_fwd(const void* d_x, const float* d_gamma, const float* d_beta, float* d_moving_mean,
                                float* d_moving_var, int64_t M, int C, float eps, float momentum, int act,
                                const void* d_res, void* d_y, float* d_save, void* d_workspace, size_t workspace_bytes,
                                ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_x); SSD_REQUIRE_PTR(d_gamma); SSD_REQUIRE_PTR(d_beta); SSD_REQUIRE_PTR(d_y);
    SSD_REQUIRE_PTR(d_save); SSD_REQUIRE_PTR(d_workspace);
    SSD_REQUIRE(M >= 1 && C >= 8 && C % 8 == 0 && C <= 8 * kRedThreads && act >= SSD_ACT_NONE && act <= SSD_ACT_RELU6,
                SSD_ERR_SHAPE, "ssd_bn_train_fwd: bad shape M=%lld C=%d act=%d", (long long)M, C, act);
    SSD_REQUIRE(workspace_bytes >= ssd_bn_workspace_bytes(C), SSD_ERR_WORKSPACE,
                "ssd_bn_train_fwd: workspace %zu < required %zu bytes", workspace_bytes, ssd_bn_workspace_bytes(C));
    cudaStream_t st = as_stream(stream);
    const ChunkGeom g = chunk_geom(M, C);
    float* partial = static_cast<float*>(d_workspace);
    bn_stats_partial_kernel<<<g.chunks, kRedThreads, kRedThreads * 16 * sizeof(float), st>>>(
        reinterpret_cast<const uint4*>(d_x), g, partial);
    SSD_CHECK_LAUNCH("bn_stats_partial_kernel");
    const double bessel = M > 1 ? (double)M / (double)(M - 1) : 1.0;
    float* coef = partial + (size_t)kMaxChunks * 2 * C;
    bn_stats_finalize_kernel<<<ceil_div(C, 8), 256, 0, st>>>(partial, g.chunks, C, 1.0 / (double)M, bessel, eps, momentum,
                                                               d_gamma, d_beta, d_save, coef, d_moving_mean, d_moving_var);
    SSD_CHECK_LAUNCH("bn_stats_finalize_kernel");
    int64_t stride = 0;
    const int grid = ew_grid_c8(g.total, g.C8, &stride);
    bn_apply_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_x), coef, reinterpret_cast<const uint4*>(d_res),
                                           reinterpret_cast<uint4*>(d_y), g.C8, C, act, g.total, stride);
    SSD_CHECK_LAUNCH("bn_apply_kernel");
    return SSD_OK;
}

extern "C" int ssd_bn_train_bwd(const void* d_x, const void* d_dy, const float* d_gamma, const float* d_beta,
                                const float* d_save, int64_t M, int C, int act, void* d_dx, void* d_dres,
                                int accumulate_res, float* d_dgamma, float* d_dbeta, void* d_workspace,
                                size_t workspace_bytes, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_x); SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_gamma); SSD_REQUIRE_PTR(d_beta);
    SSD_REQUIRE_PTR(d_save); SSD_REQUIRE_PTR(d_dx); SSD_REQUIRE_PTR(d_workspace);
    SSD_REQUIRE(M >= 1 && C >= 8 && C % 8 == 0 && C <= 8 * kRedThreads && act >= SSD_ACT_NONE && act <= SSD_ACT_RELU6,
                SSD_ERR_SHAPE, "ssd_bn_train_bwd: bad shape M=%lld C=%d act=%d", (long long)M, C, act);
    SSD_REQUIRE(workspace_bytes >= ssd_bn_workspace_bytes(C), SSD_ERR_WORKSPACE,
                "ssd_bn_train_bwd: workspace %zu < required %zu bytes", workspace_bytes, ssd_bn_workspace_bytes(C));
    cudaStream_t st = as_stream(stream);
    const ChunkGeom g = chunk_geom(M, C);
    float* partial = static_cast<float*>(d_workspace);
    float* coef = partial + (size_t)kMaxChunks * 2 * C;
    bn_bwd_partial_kernel<<<g.chunks, kRedThreads, kRedThreads * 16 * sizeof(float), st>>>(
        reinterpret_cast<const uint4*>(d_x), reinterpret_cast<const uint4*>(d_dy), d_gamma, d_beta, d_save, C, act, g,
        partial);
    SSD_CHECK_LAUNCH("bn_bwd_partial_kernel");
    bn_bwd_finalize_kernel<<<ceil_div(C, 8), 256, 0, st>>>(partial, g.chunks, C, 1.0 / (double)M, d_dgamma, d_dbeta, coef);
    SSD_CHECK_LAUNCH("bn_bwd_finalize_kernel");
    int64_t stride = 0;
    const int grid = ew_grid_c8(g.total, g.C8, &stride);
    bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(
        reinterpret_cast<const uint4*>(d_x), reinterpret_cast<const uint4*>(d_dy), d_gamma, d_beta, d_save, coef, g.C8, C,
        act, reinterpret_cast<uint4*>(d_dx), reinterpret_cast<uint4*>(d_dres), accumulate_res, g.total, stride);
    SSD_CHECK_LAUNCH("bn_bwd_apply_kernel");
    return SSD_OK;
}

extern "C" int ssd_depthwise3x3_dgrad(const void* d_dy, const void* d_weight, void* d_dx, int B, int H, int W, int C,
                                      int Ho, int Wo, int stride, int pad_top, int pad_left, int accumulate,
                                      ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_weight); SSD_REQUIRE_PTR(d_dx);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && Ho >= 1 && Wo >= 1 && (stride == 1 || stride == 2),
                SSD_ERR_SHAPE, "ssd_depthwise3x3_dgrad: bad shape B=%d H=%d W=%d C=%d Ho=%d Wo=%d stride=%d", B, H, W, C,
                Ho, Wo, stride);
    const int64_t total = (int64_t)B * H * W * (C / 8);
    dw_dgrad_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint4*>(d_dy), reinterpret_cast<const uint4*>(d_weight), reinterpret_cast<uint4*>(d_dx), H, W,
        C / 8, Ho, Wo, stride, pad_top, pad_left, accumulate, total);
    SSD_CHECK_LAUNCH("dw_dgrad_kernel");
    return SSD_OK;
}

extern "C" int ssd_depthwise3x3_wgrad(const void* d_x, const void* d_dy, float* d_dw, int B, int H, int W, int C, int Ho,
                                      int Wo, int stride, int pad_top, int pad_left, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_x); SSD_REQUIRE_PTR(d_dy); SSD_REQUIRE_PTR(d_dw);
    SSD_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && C <= 8 * kRedThreads && Ho >= 1 && Wo >= 1 &&
                (stride == 1 || stride == 2), SSD_ERR_SHAPE,
                "ssd_depthwise3x3_wgrad: bad shape B=%d H=%d W=%d C=%d Ho=%d Wo=%d stride=%d", B, H, W, C, Ho, Wo, stride);
    const ChunkGeom g = chunk_geom((int64_t)B * Ho * Wo, C, 4);
    dw_wgrad_kernel<<<g.chunks, kRedThreads, kRedThreads * 24 * sizeof(float), as_stream(stream)>>>(
        reinterpret_cast<const uint4*>(d_x), reinterpret_cast<const uint4*>(d_dy), H, W, Ho, Wo, stride, pad_top, pad_left,
        g, d_dw);
    SSD_CHECK_LAUNCH("dw_wgrad_kernel");
    return SSD_OK;
}
