// A CHAIN of small convolutions as ONE launch: the tail of the SSD networks -- models/ssd_mobilenet_v2.py:33-41
// (extra2_1 ... extra4_2), models/ssd_vgg16.py:108-113 (conv9_1 ... conv11_2) -- together with the multibox heads of
// the feature maps the tail produces (models/header.py:68-85).  These layers see 5x5 ... 1x1 maps: a few MFLOP each,
// so as separate launches every one of them costs a launch + pipeline fill (10-20 us) for < 1 us of work, and they
// sit on the critical path one behind the other.
//
// Images are independent through the whole tail, so a thread-block CLUSTER of 8 CTAs owns `ipc` images from the first
// layer to the last: every CTA computes its share of the OUTPUT CHANNELS of a layer (8-channel tiles, round-robin over
// the cluster ranks) for all pixels of the cluster's images, writes it to the layer's output tensor in global memory
// (L2), and a cluster barrier (release / acquire) makes it visible to the peers, which stage the whole activation
// back into shared memory for the next layer.  Layers of one "phase" (e.g. a head and the next extra layer, which both
// read the same map) run between two barriers.  Per layer and CTA:
//   * the input map of the cluster's images is staged in shared memory with cp.async (pixel pitch Cin*2 + 16 bytes:
//     conflict-free fragment loads);
//   * the CTA's weight rows (taps that fall on padding for every output pixel are skipped) are fetched whole with
//     cp.async -- the NEXT layer's already while this layer's epilogue and the cluster barrier run;
//   * a warp owns an (m-tile, k-slice) unit and ALL n-tiles of the CTA, so the A fragments are read from shared memory
//     once per k-step (fragment loads, not FLOPs, bound these layers); idle warps split K, partial sums are combined in a
//     fixed order;
//   * mma.sync.m16n8k16 (fp16 in, fp32 accumulate): the problem is latency-, not throughput-bound (M = 4 ... 100 rows),
//     tcgen05's 128-row tiles would be > 90 % padding here;
//   * epilogue: bias + activation -> fp16 NHWC, or the heads' fp32 scatter into the concatenated [B,N,L] / [B,N,4].

#include "common.cuh"

#include <string.h>

namespace ssd {

constexpr int CH_MAX_LAYERS = 16;
constexpr int CH_CLUSTER = 8;
constexpr int CH_THREADS = 512;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr int CH_MAX_N = 4;                         // 8-channel tiles of one CTA per layer (32 weight rows)
constexpr int CH_MAX_ROWS = CH_WARPS * 16;          // output pixels of a cluster per layer (m-tiles <= warps)
constexpr int CH_PART_BYTES = CH_WARPS * CH_MAX_N * 128 * 4;        // partial accumulators of all units (32 KB)

struct ChainLayer {
    const __half* in; const __half* w; const float* bias; void* out0; void* out1;
    long long img0, pix0, img1, pix1;               // output strides in elements (ssd_conv_desc)
    int H, W, Cin, Ho, Wo, Cout, k, stride, pad_t, pad_l, act, out_f32, split, phase;
    int tapmask, n_live;                            // bit tap: the tap touches the image for some output pixel; their count
    int pitch;                                      // shared-memory bytes per input pixel (Cin*2 + 16)
    int wpitch;                                     // shared-memory bytes per weight row (n_live*Cin*2 + 16)
    int w_off;                                      // shared-memory offset of the weight rows (they end at the partial sums)
};
struct ChainParams {
    int n_layers, B, ipc, part_off;                 // shared memory: staged input map ... weight rows | partial sums
    unsigned long long* trace;                      // debug (ssd_debug_trace): layer timeline of CTA 0, else nullptr
    ChainLayer L[CH_MAX_LAYERS];
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// this CTA's weight rows of layer L (8-channel tiles rank, rank + 8, ...; live taps only, packed) -> shared memory;
// 16 threads per row, no divisions
__device__ __forceinline__ void chain_issue_weights(const ChainLayer& L, int rank, unsigned char* smem) {
    const int n_tiles = (L.Cout + 7) >> 3;
    const int my_n = rank < n_tiles ? (n_tiles - rank + CH_CLUSTER - 1) / CH_CLUSTER : 0;
    const int row = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    if (row >= my_n * 8) return;
    const int Cin = L.Cin, kk = L.k * L.k, cpr = Cin >> 3;
    const int n = min((rank + CH_CLUSTER * (row >> 3)) * 8 + (row & 7), L.Cout - 1);
    const __half* src = L.w + (size_t)n * kk * Cin;
    unsigned char* dst = smem + L.w_off + (size_t)row * L.wpitch;
    const int mask = L.tapmask;
    int lt = 0;
    for (int tap = 0; tap < kk; ++tap) {
        if (!((mask >> tap) & 1)) continue;
        for (int ch = l16; ch < cpr; ch += 16) cp_async16(dst + ((size_t)lt * Cin + ch * 8) * 2, src + (size_t)tap * Cin + ch * 8);
        ++lt;
    }
}

__global__ void __cluster_dims__(CH_CLUSTER, 1, 1) __launch_bounds__(CH_THREADS, 1)
conv_chain_kernel(const __grid_constant__ ChainParams p) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    __shared__ int sRow[CH_MAX_ROWS];                               // output row -> (input image base << 16 | oy << 8 | ox)
    unsigned char* sAct = ch_smem;
    float* sPart = reinterpret_cast<float*>(ch_smem + p.part_off);
    const uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(ch_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int rank = (int)cluster_rank();
    const int cid = (int)blockIdx.x / CH_CLUSTER;
    const int b0 = cid * p.ipc, nb = min(p.ipc, p.B - b0);          // images of this cluster (nb >= 1 by grid size)

    chain_issue_weights(p.L[0], rank, ch_smem);                     // weights never depend on the peers
    cp_async_commit();
    int prev_phase = p.L[0].phase;
    int tslot = 0;
    for (int l = 0; l < p.n_layers; ++l) {
        // the layer's geometry in registers (the parameter block is indexed dynamically)
        const ChainLayer& Lp = p.L[l];
        const int H = Lp.H, W = Lp.W, Cin = Lp.Cin, Ho = Lp.Ho, Wo = Lp.Wo, Cout = Lp.Cout, ksz = Lp.k, stride = Lp.stride;
        const int pad_t = Lp.pad_t, pad_l = Lp.pad_l, tapmask = Lp.tapmask, n_live = Lp.n_live, pitch = Lp.pitch, wpitch = Lp.wpitch;
        if (tid == 0) trace_stamp(p.trace, 0, tslot, 1);
        if (Lp.phase != prev_phase) { cluster_sync_all(); prev_phase = Lp.phase; }   // the peers' outputs are visible now
        if (tid == 0) trace_stamp(p.trace, 0, tslot, 2);

        // ---- stage the input map of the cluster's images (cp.async.cg reads L2: the data came from peer CTAs) ----
        {
            const int npix_in = nb * H * W, cpp = Cin >> 3;         // 16-byte chunks per pixel
            const __half* src = Lp.in + (size_t)b0 * H * W * Cin;
            const int l16 = tid & 15;
            for (int pix = tid >> 4; pix < npix_in; pix += CH_THREADS / 16)
                for (int ch = l16; ch < cpp; ch += 16)
                    cp_async16(sAct + (size_t)pix * pitch + ch * 16, src + (size_t)pix * Cin + ch * 8);
            cp_async_commit();
        }
        const int HoWo = Ho * Wo, M = nb * HoWo, m_tiles = (M + 15) >> 4;
        const int n_tiles = (Cout + 7) >> 3;
        const int my_n = rank < n_tiles ? (n_tiles - rank + CH_CLUSTER - 1) / CH_CLUSTER : 0;
        // bias of the (up to two) output pairs this warp finishes: requested now, needed in the epilogue
        float bias_q[2][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int pi = warp + CH_WARPS * q;
            const int nj = my_n > 0 ? pi % my_n : 0;
            const int n0 = (rank + CH_CLUSTER * nj) * 8 + 2 * t;
            bias_q[q][0] = (Lp.bias && pi < m_tiles * my_n && n0 < Cout) ? __ldg(Lp.bias + n0) : 0.0f;
            bias_q[q][1] = (Lp.bias && pi < m_tiles * my_n && n0 + 1 < Cout) ? __ldg(Lp.bias + n0 + 1) : 0.0f;
        }
        for (int r = tid; r < m_tiles * 16; r += CH_THREADS) {
            const int rr = r < M ? r : 0;
            const int img = rr / HoWo, pp = rr - img * HoWo, oy = pp / Wo, ox = pp - oy * Wo;
            sRow[r] = r < M ? ((img * H * W) << 16) | (oy << 8) | ox : -1;
        }
        cp_async_wait_all();
        __syncthreads();                                            // input map + this layer's weights are in shared memory
        if (tid == 0) trace_stamp(p.trace, 0, tslot, 3);

        // ---- units: (m-tile, k-slice); a unit computes ALL n-tiles of this CTA, so the A fragments are loaded once ----
        const int S = max(1, CH_WARPS / m_tiles);                   // k-slices per m-tile (idle warps split K)
        const int spt = Cin >> 4;                                   // k-steps (16 channels) per tap
        const int T = n_live * spt;                                 // k-steps of the layer
        if (warp < m_tiles * S && my_n > 0) {
            const int mt = warp / S, sl = warp - mt * S;
            float acc[CH_MAX_N][4];
#pragma unroll
            for (int j = 0; j < CH_MAX_N; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
            const int ri0 = sRow[mt * 16 + g], ri1 = sRow[mt * 16 + g + 8];
            const uint32_t wrow = smem32 + (uint32_t)Lp.w_off + (uint32_t)g * (uint32_t)wpitch + (uint32_t)(2 * t) * 2u;
            const uint32_t wtile = 8u * (uint32_t)wpitch;           // next n-tile of this CTA
            const int k_lo = (T * sl) / S, k_hi = (T * (sl + 1)) / S;
            int lt = 0;
            for (int tap = 0; tap < ksz * ksz; ++tap) {
                if (!((tapmask >> tap) & 1)) continue;
                const int s0 = max(k_lo, lt * spt), s1 = min(k_hi, (lt + 1) * spt);
                if (s0 < s1) {
                    const int ky = tap / ksz, kx = tap - ky * ksz;
                    uint32_t aoff[2], am[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int ri = h ? ri1 : ri0;
                        const int iy = ((ri >> 8) & 0xff) * stride - pad_t + ky, ix = (ri & 0xff) * stride - pad_l + kx;
                        const bool ok = ri >= 0 && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
                        am[h] = ok ? 0xffffffffu : 0u;              // rows on padding read pixel 0 and are masked
                        aoff[h] = smem32 + (uint32_t)(ok ? (ri >> 16) + iy * W + ix : 0) * (uint32_t)pitch + (uint32_t)(2 * t) * 2u;
                    }
                    const uint32_t wtap = wrow + (uint32_t)(lt * Cin) * 2u;
                    for (int ks = s0 - lt * spt; ks < s1 - lt * spt; ++ks) {
                        const uint32_t o = (uint32_t)ks * 32u;
                        const uint32_t a0 = lds32(aoff[0] + o) & am[0], a2 = lds32(aoff[0] + o + 16) & am[0];
                        const uint32_t a1 = lds32(aoff[1] + o) & am[1], a3 = lds32(aoff[1] + o + 16) & am[1];
#pragma unroll
                        for (int j = 0; j < CH_MAX_N; ++j) {
                            if (j < my_n) {                         // warp-uniform
                                const uint32_t b0r = lds32(wtap + j * wtile + o), b1r = lds32(wtap + j * wtile + o + 16);
                                mma16816(acc[j], a0, a1, a2, a3, b0r, b1r);
                            }
                        }
                    }
                }
                ++lt;
            }
#pragma unroll
            for (int j = 0; j < CH_MAX_N; ++j)
                if (j < my_n)
                    reinterpret_cast<float4*>(sPart + (size_t)(warp * CH_MAX_N + j) * 128)[lane] =
                        make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        }
        __syncthreads();                                            // everyone is done with the staged map and weights
        if (tid == 0) trace_stamp(p.trace, 0, tslot, 4);
        if (l + 1 < p.n_layers) chain_issue_weights(p.L[l + 1], rank, ch_smem);      // in flight across the epilogue + barrier
        cp_async_commit();

        // ---- epilogue: the (m-tile, n-tile) pairs over the warps; k-slices summed in a fixed order; bias + activation ->
        //      fp16 NHWC, or the heads' fp32 scatter ----
        const int act = Lp.act, out_f32 = Lp.out_f32, split = Lp.split;
        const long long img0 = Lp.img0, pix0 = Lp.pix0, img1 = Lp.img1, pix1 = Lp.pix1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int pi = warp + CH_WARPS * q;
            if (pi >= m_tiles * my_n) break;
            const int mt = pi / my_n, nj = pi - mt * my_n;
            float v4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int sl = 0; sl < S; ++sl) {
                const float4 o = reinterpret_cast<const float4*>(sPart + (size_t)((mt * S + sl) * CH_MAX_N + nj) * 128)[lane];
                v4[0] += o.x; v4[1] += o.y; v4[2] += o.z; v4[3] += o.w;
            }
            const int n0 = (rank + CH_CLUSTER * nj) * 8 + 2 * t;    // channels n0, n0 + 1 of this thread
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = mt * 16 + g + 8 * h;
                if (r >= M) continue;
                const int img = r / HoWo, pp = r - img * HoWo;
                float v[2] = {v4[2 * h], v4[2 * h + 1]};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = n0 + e;
                    if (n < Cout) {
                        v[e] += bias_q[q][e];
                        if (act == SSD_ACT_RELU) v[e] = fmaxf(v[e], 0.0f);
                        else if (act == SSD_ACT_RELU6) v[e] = fminf(fmaxf(v[e], 0.0f), 6.0f);
                    }
                }
                const long long b = b0 + img;
                if (out_f32) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int n = n0 + e;
                        if (n >= Cout) continue;
                        if (n < split) static_cast<float*>(Lp.out0)[b * img0 + (long long)pp * pix0 + n] = v[e];
                        else           static_cast<float*>(Lp.out1)[b * img1 + (long long)pp * pix1 + (n - split)] = v[e];
                    }
                } else {
                    __half* o = static_cast<__half*>(Lp.out0) + b * img0 + (long long)pp * pix0 + n0;
                    if (n0 + 1 < Cout) *reinterpret_cast<__half2*>(o) = __floats2half2_rn(v[0], v[1]);
                    else if (n0 < Cout) o[0] = __float2half_rn(v[0]);
                }
            }
        }
    }
    cp_async_wait_all();
    if (tid == 0) trace_stamp(p.trace, 0, tslot, 5);
}

// taps that touch the image for at least one output pixel (a tap that is padding everywhere contributes nothing)
static int live_taps(const ssd_conv_desc& d) {
    int mask = 0;
    for (int ky = 0; ky < d.KH; ++ky)
        for (int kx = 0; kx < d.KW; ++kx) {
            bool any_y = false, any_x = false;
            for (int oy = 0; oy < d.Ho; ++oy) { const int iy = oy * d.stride - d.pad_top + ky; any_y |= iy >= 0 && iy < d.H; }
            for (int ox = 0; ox < d.Wo; ++ox) { const int ix = ox * d.stride - d.pad_left + kx; any_x |= ix >= 0 && ix < d.W; }
            if (any_y && any_x) mask |= 1 << (ky * d.KW + kx);
        }
    return mask;
}

// fills `p` (ipc, shared-memory layout, per-layer fields); false: the chain does not fit this kernel.
// max_clusters: how many clusters the device keeps resident at once (0: unknown) -- a second wave would double the time,
// so the smallest images-per-cluster count whose cluster count still fits is taken.
static bool chain_plan(const ssd_conv_desc* descs, const int32_t* phase, int n, int max_clusters, ChainParams* pp, size_t* smem_out) {
    if (n < 1 || n > CH_MAX_LAYERS) return false;
    ChainParams& p = *pp;
    const int B = descs[0].B;
    for (int i = 0; i < n; ++i) {
        const ssd_conv_desc& d = descs[i];
        if (d.B != B || d.KH != d.KW || (d.KH != 1 && d.KH != 3) || d.dilation != 1 || (d.stride != 1 && d.stride != 2) ||
            d.Cin % 128 != 0 || d.residual != nullptr || d.Cout < 1 || (i > 0 && phase[i] < phase[i - 1]) ||
            d.Ho > 255 || d.Wo > 255 ||
            (!d.out_f32 && (d.split != d.Cout || d.pix_stride0 % 2 != 0 || d.img_stride0 % 2 != 0)) ||
            (reinterpret_cast<uintptr_t>(d.in) & 15) || (reinterpret_cast<uintptr_t>(d.weight) & 15))
            return false;
    }
    const size_t budget = 225 * 1024;                               // 227 KB per CTA minus the static row table
    int best_ipc = 0;
    size_t best_need = 0;
    for (int ipc = 1; ipc <= min(8, B); ++ipc) {
        size_t need = 0;                                            // max over the layers of (input map + weight rows)
        bool ok = true;
        for (int i = 0; i < n && ok; ++i) {
            const ssd_conv_desc& d = descs[i];
            const size_t act = align_up((size_t)ipc * d.H * d.W * (d.Cin * 2 + 16), 128);
            const int m_tiles = (ipc * d.Ho * d.Wo + 15) / 16, n_tiles = (d.Cout + 7) / 8;
            const int my_n = (n_tiles + CH_CLUSTER - 1) / CH_CLUSTER;
            const int live = __builtin_popcount(live_taps(d));
            const size_t wb = align_up((size_t)my_n * 8 * ((size_t)live * d.Cin * 2 + 16), 128);
            need = max(need, act + wb);
            ok = my_n <= CH_MAX_N && m_tiles <= CH_WARPS && m_tiles * my_n <= 2 * CH_WARPS && (size_t)ipc * d.H * d.W < 32768;
        }
        if (!ok || need + CH_PART_BYTES > budget) break;            // larger ipc only needs more
        best_ipc = ipc; best_need = need;
        if (max_clusters <= 0 ? (size_t)((B + ipc - 1) / ipc) * CH_CLUSTER <= (size_t)sm_count() - 2 * CH_CLUSTER
                              : (B + ipc - 1) / ipc <= max_clusters)
            break;                                                  // all clusters resident in one wave
    }
    if (!best_ipc) return false;
    const int ipc = best_ipc;
    p.n_layers = n; p.B = B; p.ipc = ipc; p.part_off = (int)best_need;
    for (int i = 0; i < n; ++i) {
        const ssd_conv_desc& d = descs[i];
        ChainLayer& L = p.L[i];
        L.in = static_cast<const __half*>(d.in); L.w = static_cast<const __half*>(d.weight); L.bias = d.bias;
        L.out0 = d.out0; L.out1 = d.out1;
        L.img0 = d.img_stride0; L.pix0 = d.pix_stride0; L.img1 = d.img_stride1; L.pix1 = d.pix_stride1;
        L.H = d.H; L.W = d.W; L.Cin = d.Cin; L.Ho = d.Ho; L.Wo = d.Wo; L.Cout = d.Cout; L.k = d.KH; L.stride = d.stride;
        L.pad_t = d.pad_top; L.pad_l = d.pad_left; L.act = d.act; L.out_f32 = d.out_f32; L.split = d.split; L.phase = phase[i];
        L.tapmask = live_taps(d);
        L.n_live = __builtin_popcount(L.tapmask);
        L.pitch = d.Cin * 2 + 16;
        L.wpitch = L.n_live * d.Cin * 2 + 16;
        const int my_n = ((d.Cout + 7) / 8 + CH_CLUSTER - 1) / CH_CLUSTER;
        L.w_off = (int)(best_need - align_up((size_t)my_n * 8 * L.wpitch, 128));     // weight rows end where the partial sums begin
    }
    *smem_out = best_need + CH_PART_BYTES;
    return true;
}

}  // namespace ssd

using namespace ssd;

// clusters of this kernel the device keeps resident at once with `smem` bytes of dynamic shared memory (0: unknown)
static bool chain_ensure_attr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    static thread_local int attr_dev = -1;
    if (attr_dev != dev) {
        if (cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_dev = dev;
    }
    return true;
}
static int chain_max_clusters(size_t smem) {
    if (!chain_ensure_attr()) return 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(CH_CLUSTER * 64); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CH_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_chain_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int ssd_conv_chain_supported(const ssd_conv_desc* h_descs, const int32_t* h_phase, int n_layers) {
    if (!h_descs || !h_phase) return 0;
    ChainParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0;
    return chain_plan(h_descs, h_phase, n_layers, 0, &p, &smem) ? 1 : 0;
}

// 0: unsupported; otherwise how many waves of clusters the launch needs on the current device (1 = every cluster is
// resident at once; the engine fuses a tail only then: a second wave would double the latency the fusion removes)
extern "C" int ssd_conv_chain_waves(const ssd_conv_desc* h_descs, const int32_t* h_phase, int n_layers) {
    if (!h_descs || !h_phase) return 0;
    ChainParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0;
    static thread_local int max_clusters = -1;
    if (max_clusters < 0) max_clusters = chain_max_clusters(225 * 1024);
    if (!chain_plan(h_descs, h_phase, n_layers, max_clusters, &p, &smem)) return 0;
    const int clusters = (p.B + p.ipc - 1) / p.ipc;
    return max_clusters > 0 ? (clusters + max_clusters - 1) / max_clusters : 1;
}

extern "C" int ssd_conv_chain(const ssd_conv_desc* h_descs, const int32_t* h_phase, int n_layers, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(h_descs); SSD_REQUIRE_PTR(h_phase);
    SSD_REQUIRE(n_layers >= 1 && n_layers <= CH_MAX_LAYERS, SSD_ERR_SHAPE, "ssd_conv_chain: n_layers=%d (1..%d)", n_layers, CH_MAX_LAYERS);
    for (int i = 0; i < n_layers; ++i) {
        SSD_REQUIRE_PTR(h_descs[i].in); SSD_REQUIRE_PTR(h_descs[i].weight); SSD_REQUIRE_PTR(h_descs[i].out0);
        SSD_REQUIRE(h_descs[i].split == h_descs[i].Cout || h_descs[i].out1 != nullptr, SSD_ERR_NULL, "ssd_conv_chain: layer %d: out1 is NULL", i);
    }
    ChainParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0;
    static thread_local int max_clusters = -1;                     // measured once per thread with the largest footprint
    if (max_clusters < 0) max_clusters = chain_max_clusters(225 * 1024);
    SSD_REQUIRE(chain_ensure_attr(), SSD_ERR_UNSUPPORTED, "ssd_conv_chain: cannot raise the dynamic shared-memory limit");
    SSD_REQUIRE(chain_plan(h_descs, h_phase, n_layers, max_clusters, &p, &smem), SSD_ERR_UNSUPPORTED,
                "ssd_conv_chain: unsupported chain (k 1|3, dilation 1, stride 1|2, Cin %% 128, no residual, one batch size, "
                "non-decreasing phases, small maps)");
    const int clusters = (p.B + p.ipc - 1) / p.ipc;
    p.trace = debug_trace_buffer();
    if (p.trace)
        fprintf(stderr, "ssd_conv_chain: %d layers, ipc=%d, %d clusters (max co-resident %d), smem=%zu\n", p.n_layers, p.ipc, clusters,
                max_clusters, smem);
    conv_chain_kernel<<<clusters * CH_CLUSTER, CH_THREADS, smem, as_stream(stream)>>>(p);
    SSD_CHECK_LAUNCH("conv_chain_kernel");
    return SSD_OK;
}
