// SSD loss (ssd_loss.py:26-91) for sm_100a: Huber on positives + categorical
// cross-entropy with hard-negative mining, forward and backward.
//
// Three HBM-bound passes, none of which sorts:
//   1. loss_anchor_kernel  -- full grid, one thread per (image, anchor); the
//      [*, L] rows are staged through shared memory with 16-byte loads; emits
//      per-anchor CE, masked CE, masked Huber and flag bytes (13 B/anchor).
//   2. loss_select_kernel  -- one CTA per image: the reference's
//      argsort(argsort(masked, DESC)) < n_neg  (:79-81) is evaluated as an
//      exact radix-select of the n_neg-th largest masked loss plus an
//      index-ordered tie break, then deterministic block reductions.
//   3. loss_bwd_kernel     -- full grid, gradient w.r.t. deltas and logits.

#include "common.cuh"

namespace ssd {

constexpr int kRowThreads = 128;      // anchors per CTA in the row-staged kernels
constexpr int kSelThreads = 1024;

struct LossWs {            // views into the caller's workspace
    float*   ce;           // [B,N] per-anchor cross entropy
    float*   masked;       // [B,N] ce * y[...,0]           (ssd_loss.py:78)
    float*   hub;          // [B,N] pos_loc * huber_sum4    (ssd_loss.py:50)
    uint8_t* flags;        // [B,N] bit0 pos_loc (:46) bit1 pos_conf (:72)
    uint8_t* fmask;        // [B,N] final_mask = pos + neg in {0,1,2} (:84)
    float*   npos_loc;     // [B]   divisor after the ==0 -> 1 rule (:51-55)
    float*   npos_conf;    // [B]   (:86-90)
};

static size_t loss_ws_layout(int B, int N, LossWs* w, void* base) {
    size_t bn = (size_t)B * N, off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_ce = take(bn * 4), o_m = take(bn * 4), o_h = take(bn * 4), o_f = take(bn), o_fm = take(bn);
    size_t o_nl = take((size_t)B * 4), o_nc = take((size_t)B * 4);
    if (w) {
        char* p = static_cast<char*>(base);
        w->ce = (float*)(p + o_ce); w->masked = (float*)(p + o_m); w->hub = (float*)(p + o_h);
        w->flags = (uint8_t*)(p + o_f); w->fmask = (uint8_t*)(p + o_fm);
        w->npos_loc = (float*)(p + o_nl); w->npos_conf = (float*)(p + o_nc);
    }
    return off;
}

// ------------------------------------------------------------------ pass 1 --
template <bool FROM_LOGITS, int LT>                    // LT: compile-time label count (0 = runtime L)
__global__ void __launch_bounds__(kRowThreads)
loss_anchor_kernel(const float4* __restrict__ act_d, const float4* __restrict__ pred_d,
                   const float* __restrict__ act_l, const float* __restrict__ pred_l,
                   int N, int L_rt, LossWs w) {
    const int L = LT ? LT : L_rt;                         // the per-row loops unroll completely for the VOC label count
    extern __shared__ __align__(16) float s_rows[];   // [cnt*L + 4] labels, [cnt*L + 4] predictions
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kRowThreads;
    const int cnt = min(kRowThreads, N - n0);
    const size_t row0 = (size_t)b * N + n0;
    const int region = kRowThreads * L + 4;               // floats per staging region (multiple of 4)
    float* s_y = s_rows;
    float* s_p = s_rows + region;
    const bool do_conf = act_l != nullptr;
    const bool mine = (int)threadIdx.x < cnt;
    const size_t i = row0 + threadIdx.x;
    // every global read of the CTA is requested before anything is consumed
    if (do_conf) {
        s_y = stage_rows_in(act_l + row0 * L, cnt * L, s_rows);
        s_p = stage_rows_in(pred_l + row0 * L, cnt * L, s_rows + region);
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), p4 = a;
    if (act_d != nullptr && mine) { a = __ldcs(act_d + i); p4 = __ldcs(pred_d + i); }
    if (do_conf) {
        stage_rows_wait();
        __syncthreads();
    }
    if (!mine) return;
    uint8_t flag = 0;

    if (act_d != nullptr) {                           // ssd_loss.py:36-50
        float av[4] = {a.x, a.y, a.z, a.w}, pv[4] = {p4.x, p4.y, p4.z, p4.w};
        float hsum = 0.0f;
        bool pos = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float e = fabsf(fsub(pv[k], av[k]));
            float q = fminf(e, 1.0f);
            float lin = fsub(e, q);
            hsum = fadd(hsum, fadd(fmul(0.5f, fmul(q, q)), fmul(1.0f, lin)));
            pos |= (av[k] != 0.0f);
        }
        w.hub[i] = pos ? hsum : 0.0f;
        flag |= pos ? 1 : 0;
    }
    if (do_conf) {                                    // ssd_loss.py:69-78
        const float* y = s_y + (size_t)threadIdx.x * L;
        const float* p = s_p + (size_t)threadIdx.x * L;
        // label scan: foreground flag (:72), number of non-zero entries and the last one.  Targets
        // are one-hot in practice, so the sum over classes (:70) has at most one non-zero term.
        uint32_t fg_bits = 0;
        int nnz = 0, last = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const uint32_t bits = __float_as_uint(y[l]) & 0x7fffffffu;     // +-0 -> 0
            if (l > 0) fg_bits |= bits;
            if (bits) { ++nnz; last = l; }
        }
        float ce = 0.0f;
        if (FROM_LOGITS) {
            float m = p[0];
#pragma unroll
            for (int l = 1; l < L; ++l) m = fmaxf(m, p[l]);
            float s = 0.0f;
#pragma unroll
            for (int l = 0; l < L; ++l) s = fadd(s, expf(fsub(p[l], m)));
            const float lse = fadd(logf(s), m);
            if (nnz == 1) {
                ce = fadd(0.0f, fmul(y[last], fsub(p[last], lse)));
            } else if (nnz > 1) {
                for (int l = 0; l < L; ++l) {
                    float yl = y[l];
                    if (yl != 0.0f) ce = fadd(ce, fmul(yl, fsub(p[l], lse)));
                }
            }
        } else {
            float s = 0.0f;
#pragma unroll
            for (int l = 0; l < L; ++l) s = fadd(s, p[l]);
            if (nnz == 1) {
                float q = fminf(fmaxf(fdiv(p[last], s), 1e-7f), fsub(1.0f, 1e-7f));
                ce = fadd(0.0f, fmul(y[last], logf(q)));
            } else if (nnz > 1) {
                for (int l = 0; l < L; ++l) {
                    float yl = y[l];
                    if (yl != 0.0f) {
                        float q = fminf(fmaxf(fdiv(p[l], s), 1e-7f), fsub(1.0f, 1e-7f));
                        ce = fadd(ce, fmul(yl, logf(q)));
                    }
                }
            }
        }
        ce = -ce;
        w.ce[i] = ce;
        w.masked[i] = fmul(ce, y[0]);
        flag |= fg_bits ? 2 : 0;
    }
    w.flags[i] = flag;
}

// ------------------------------------------------------------------ pass 2 --
__device__ __forceinline__ uint32_t order_key(float f) {     // monotone float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ float block_sum(float v, float* s_red) {           // deterministic tree
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) s_red[0] = t;
    }
    __syncthreads();
    float r = s_red[0];
    __syncthreads();
    return r;
}

// Three block sums at once (deterministic tree).
__device__ float3 block_sum3(float3 v, float3* s_red3) {
    v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_red3[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float3 t = (lane < (int)(blockDim.x >> 5)) ? s_red3[lane] : make_float3(0.f, 0.f, 0.f);
        t.x = warp_sum(t.x); t.y = warp_sum(t.y); t.z = warp_sum(t.z);
        if (lane == 0) s_red3[0] = t;
    }
    __syncthreads();
    float3 r = s_red3[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kSelThreads)
loss_select_kernel(int N, float neg_pos_ratio, float alpha, bool do_loc, bool do_conf, LossWs w,
                   float* __restrict__ out_loc, float* __restrict__ out_conf) {
    extern __shared__ uint32_t s_key[];               // [N] keys (the flag bytes are re-read from L2 in the last pass:
                                                      //  4 B/anchor of shared memory lets two CTAs share an SM up to N = 28 000)
    __shared__ float s_red[32];
    __shared__ float3 s_red3[32];
    __shared__ int s_hist[256];
    __shared__ int s_scan[kSelThreads / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const size_t base = (size_t)b * N;
    const int chunk = (N + kSelThreads - 1) / kSelThreads;
    const int lo = min(N, tid * chunk), hi = min(N, lo + chunk);

    // counts (exact in float while N < 2^24, like the reference's float sums)
    float3 acc = make_float3(0.f, 0.f, 0.f);          // n_pos_loc, n_pos_conf, huber sum
    for (int i = tid; i < N; i += kSelThreads) {
        const uint8_t f = w.flags[base + i];
        acc.x += (f & 1) ? 1.0f : 0.0f;
        acc.y += (f & 2) ? 1.0f : 0.0f;
        if (do_loc) acc.z += w.hub[base + i];
        if (do_conf) s_key[i] = order_key(fadd(w.masked[base + i], 0.0f));   // -0 ranks like +0 (argsort ties)
    }
    acc = block_sum3(acc, s_red3);
    const float n_pos_loc = acc.x, n_pos_conf = acc.y;
    if (do_loc) {
        float div = (n_pos_loc == 0.0f) ? 1.0f : n_pos_loc;       // :51-55
        if (tid == 0) {
            out_loc[b] = fmul(fdiv(acc.z, div), alpha);            // :56-57
            w.npos_loc[b] = div;
        }
    }
    if (!do_conf) return;

    // :75  total_neg = int32(total_pos * ratio); rank < total_neg selects the
    // k = min(total_neg, N) largest masked losses, ties by lower index.
    int k = __float2int_rz(fmul(n_pos_conf, neg_pos_ratio));
    k = max(0, min(k, N));
    uint32_t T = 0;
    int need_eq = 0;                                  // how many keys == T are selected
    if (k > 0 && k < N) {
        if (tid == 0) { s_prefix = 0; s_remaining = k; }
        uint32_t mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kSelThreads) s_hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i0 = 0; i0 < N; i0 += kSelThreads) {
                int i = i0 + tid;
                bool in = i < N && ((s_key[i < N ? i : 0] & mask) == prefix);
                uint32_t digit = in ? ((s_key[i] >> shift) & 255u) : 256u;
                // warp-aggregated histogram update
                uint32_t peers = __match_any_sync(0xffffffffu, digit);
                if (in && lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
            }
            __syncthreads();
            if (wid == 0) {
                // bins from the top: lane l owns bins 255-8l .. 248-8l; suffix counts by shuffle scan
                const int rem = s_remaining;
                int h[8], t = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { h[q] = s_hist[255 - 8 * lane - q]; t += h[q]; }
                int incl = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                const int before = incl - t;          // keys in strictly higher bins than this lane's
                // the digit d is the highest bin with (count of keys in bins > d) + hist[d] >= rem;
                // if no bin satisfies it the digit is 0 (cannot happen while rem <= matching keys)
                const bool owner = before < rem && incl >= rem;
                const bool none = __ballot_sync(0xffffffffu, owner) == 0u;
                if (owner) {
                    int a = before, d = 255 - 8 * lane;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (a + h[q] >= rem) { d = 255 - 8 * lane - q; break; }
                        a += h[q];
                    }
                    s_remaining = rem - a;
                    s_prefix = prefix | ((uint32_t)d << shift);
                } else if (none && lane == 0) {
                    s_prefix = prefix;                // unreachable (rem never exceeds the matching keys): digit 0
                }
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        T = s_prefix;
        need_eq = s_remaining;
    }

    // index-ordered rank among the keys equal to T (each thread owns a
    // contiguous index range, so an exclusive scan of per-thread counts gives
    // the number of equal keys at lower indices)
    int my_eq = 0;
    if (k > 0 && k < N)
        for (int i = lo; i < hi; ++i) my_eq += (s_key[i] == T);
    int incl = my_eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = s_scan[lane], t2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, t2, o);
            if (lane >= o) t2 += t;
        }
        s_scan[lane] = t2 - v;                        // exclusive warp offsets
    }
    __syncthreads();
    int eq_rank = s_scan[wid] + incl - my_eq;

    float total = 0.0f;
    for (int i = lo; i < hi; ++i) {
        int neg;
        if (k >= N) neg = 1;
        else if (k == 0) neg = 0;
        else {
            uint32_t key = s_key[i];
            neg = key > T ? 1 : (key == T ? (eq_rank++ < need_eq ? 1 : 0) : 0);
        }
        int fm = neg + ((w.flags[base + i] & 2) ? 1 : 0);          // :84
        w.fmask[base + i] = (uint8_t)fm;
        if (fm) total = fadd(total, fmul((float)fm, w.ce[base + i]));    // :85
    }
    total = block_sum(total, s_red);
    if (tid == 0) {
        float div = (n_pos_conf == 0.0f) ? 1.0f : n_pos_conf;     // :86-90
        out_conf[b] = fdiv(total, div);                            // :91
        w.npos_conf[b] = div;
    }
}

// ------------------------------------------------------------------ pass 3 --
__global__ void __launch_bounds__(kRowThreads)
loss_bwd_kernel(const float4* __restrict__ act_d, const float4* __restrict__ pred_d,
                const float* __restrict__ act_l, const float* __restrict__ logits,
                int N, int L, float alpha, float gscale, LossWs w,
                float4* __restrict__ g_d, float* __restrict__ g_z) {
    extern __shared__ __align__(16) float s_rows[];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kRowThreads;
    const int cnt = min(kRowThreads, N - n0);
    const size_t row0 = (size_t)b * N + n0;
    const int region = kRowThreads * L + 4;
    float* s_y = s_rows;
    float* s_z = s_rows + region;
    const bool do_conf = g_z != nullptr;
    if (do_conf) {
        s_y = stage_rows_in(act_l + row0 * L, cnt * L, s_rows);
        s_z = stage_rows_in(logits + row0 * L, cnt * L, s_rows + region);   // g_z + row0*L is misaligned like logits
        stage_rows_wait();
        __syncthreads();
    }
    if ((int)threadIdx.x < cnt) {
        const size_t i = row0 + threadIdx.x;
        if (g_d != nullptr) {
            float4 a = __ldcs(act_d + i), p = __ldcs(pred_d + i);
            float c = (w.flags[i] & 1) ? alpha * gscale / w.npos_loc[b] : 0.0f;
            float4 g;
            g.x = c * fminf(fmaxf(p.x - a.x, -1.0f), 1.0f);
            g.y = c * fminf(fmaxf(p.y - a.y, -1.0f), 1.0f);
            g.z = c * fminf(fmaxf(p.z - a.z, -1.0f), 1.0f);
            g.w = c * fminf(fmaxf(p.w - a.w, -1.0f), 1.0f);
            __stcs(g_d + i, g);
        }
        if (do_conf) {
            const float* y = s_y + (size_t)threadIdx.x * L;
            float* z = s_z + (size_t)threadIdx.x * L;
            float c = (float)w.fmask[i] * gscale / w.npos_conf[b];
            float m = z[0], ysum = 0.0f;
            for (int l = 1; l < L; ++l) m = fmaxf(m, z[l]);
            float s = 0.0f;
            for (int l = 0; l < L; ++l) { s += expf(z[l] - m); ysum += y[l]; }
            float inv = 1.0f / s;
            for (int l = 0; l < L; ++l) z[l] = c * (expf(z[l] - m) * inv * ysum - y[l]);
        }
    }
    if (do_conf) {
        __syncthreads();
        float* dst = g_z + row0 * L;
        if (stage_rows_ptr(dst, s_rows + region) == s_z) {
            stage_rows_out(dst, cnt * L, s_z);
        } else {
            for (int e = threadIdx.x; e < cnt * L; e += blockDim.x) __stcs(dst + e, s_z[e]);
        }
    }
}

}  // namespace ssd

using namespace ssd;

extern "C" size_t ssd_loss_workspace_bytes(int B, int N, int L) {
    (void)L;
    if (B <= 0 || N <= 0) return 256;
    return loss_ws_layout(B, N, nullptr, nullptr);
}

extern "C" int ssd_loss_fwd(const float* d_actual_deltas, const float* d_pred_deltas,
                            const float* d_actual_labels, const float* d_pred_labels,
                            int B, int N, int L, float neg_pos_ratio, float loc_loss_alpha, int from_logits,
                            float* d_loc_loss, float* d_conf_loss,
                            void* d_workspace, size_t workspace_bytes, ssd_stream_t stream) {
    const bool do_loc = d_loc_loss != nullptr, do_conf = d_conf_loss != nullptr;
    SSD_REQUIRE(do_loc || do_conf, SSD_ERR_NULL, "ssd_loss_fwd: both outputs are NULL");
    if (do_loc) { SSD_REQUIRE_PTR(d_actual_deltas); SSD_REQUIRE_PTR(d_pred_deltas); }
    if (do_conf) { SSD_REQUIRE_PTR(d_actual_labels); SSD_REQUIRE_PTR(d_pred_labels); }
    SSD_REQUIRE(B >= 0 && N >= 0 && L >= 1 && B <= 65535 && N < (1 << 24), SSD_ERR_SHAPE,
                "ssd_loss_fwd: bad shape B=%d N=%d L=%d", B, N, L);
    if (B == 0) return SSD_OK;
    SSD_REQUIRE(N >= 1, SSD_ERR_SHAPE, "ssd_loss_fwd: N must be >= 1");
    SSD_REQUIRE_PTR(d_workspace);
    size_t need = loss_ws_layout(B, N, nullptr, nullptr);
    SSD_REQUIRE(workspace_bytes >= need, SSD_ERR_WORKSPACE,
                "ssd_loss_fwd: workspace %zu < required %zu bytes", workspace_bytes, need);
    LossWs w;
    loss_ws_layout(B, N, &w, d_workspace);
    cudaStream_t st = as_stream(stream);

    size_t smem1 = do_conf ? (size_t)2 * (kRowThreads * L + 4) * sizeof(float) : 0;
    SSD_REQUIRE(smem1 <= 200 * 1024, SSD_ERR_UNSUPPORTED, "ssd_loss_fwd: L=%d too large for row staging", L);
    dim3 grid1(ceil_div(N, kRowThreads), B);
    auto launch1 = [&](auto kern) {
        if (smem1 > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        kern<<<grid1, kRowThreads, smem1, st>>>(
            do_loc ? reinterpret_cast<const float4*>(d_actual_deltas) : nullptr,
            reinterpret_cast<const float4*>(d_pred_deltas),
            do_conf ? d_actual_labels : nullptr, d_pred_labels, N, L, w);
    };
    if (L == 21) { if (from_logits) launch1(loss_anchor_kernel<true, 21>); else launch1(loss_anchor_kernel<false, 21>); }
    else         { if (from_logits) launch1(loss_anchor_kernel<true, 0>); else launch1(loss_anchor_kernel<false, 0>); }
    SSD_CHECK_LAUNCH("loss_anchor_kernel");

    size_t smem2 = (size_t)N * sizeof(uint32_t) + 16;
    SSD_REQUIRE(smem2 <= 200 * 1024, SSD_ERR_UNSUPPORTED,
                "ssd_loss_fwd: N=%d anchors exceed the per-image shared-memory select (max 51000)", N);
    if (smem2 > 40 * 1024)
        cudaFuncSetAttribute(loss_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    loss_select_kernel<<<B, kSelThreads, smem2, st>>>(N, neg_pos_ratio, loc_loss_alpha, do_loc, do_conf, w,
                                                      d_loc_loss, d_conf_loss);
    SSD_CHECK_LAUNCH("loss_select_kernel");
    return SSD_OK;
}

extern "C" int ssd_loss_bwd(const float* d_actual_deltas, const float* d_pred_deltas,
                            const float* d_actual_labels, const float* d_pred_logits,
                            int B, int N, int L, float loc_loss_alpha, float grad_scale,
                            float* d_grad_deltas, float* d_grad_logits,
                            const void* d_workspace, size_t workspace_bytes, ssd_stream_t stream) {
    const bool do_loc = d_grad_deltas != nullptr, do_conf = d_grad_logits != nullptr;
    SSD_REQUIRE(do_loc || do_conf, SSD_ERR_NULL, "ssd_loss_bwd: both outputs are NULL");
    if (do_loc) { SSD_REQUIRE_PTR(d_actual_deltas); SSD_REQUIRE_PTR(d_pred_deltas); }
    if (do_conf) { SSD_REQUIRE_PTR(d_actual_labels); SSD_REQUIRE_PTR(d_pred_logits); }
    SSD_REQUIRE(B >= 0 && N >= 1 && L >= 1 && B <= 65535, SSD_ERR_SHAPE,
                "ssd_loss_bwd: bad shape B=%d N=%d L=%d", B, N, L);
    if (B == 0) return SSD_OK;
    SSD_REQUIRE_PTR(d_workspace);
    size_t need = loss_ws_layout(B, N, nullptr, nullptr);
    SSD_REQUIRE(workspace_bytes >= need, SSD_ERR_WORKSPACE,
                "ssd_loss_bwd: workspace %zu < required %zu bytes", workspace_bytes, need);
    LossWs w;
    loss_ws_layout(B, N, &w, const_cast<void*>(d_workspace));
    size_t smem = do_conf ? (size_t)2 * (kRowThreads * L + 4) * sizeof(float) : 0;
    SSD_REQUIRE(smem <= 200 * 1024, SSD_ERR_UNSUPPORTED, "ssd_loss_bwd: L=%d too large for row staging", L);
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(ceil_div(N, kRowThreads), B);
    loss_bwd_kernel<<<grid, kRowThreads, smem, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(d_actual_deltas), reinterpret_cast<const float4*>(d_pred_deltas),
        d_actual_labels, d_pred_logits, N, L, loc_loss_alpha, grad_scale, w,
        reinterpret_cast<float4*>(d_grad_deltas), d_grad_logits);
    SSD_CHECK_LAUNCH("loss_bwd_kernel");
    return SSD_OK;
}
