// Box-side kernels of the tf-ssd hot path for sm_100a: prior boxes, pairwise
// IoU, fused match+encode, element-wise encode/decode.
//
// These are HBM-bound float/index kernels: the design
// rules are coalesced 16-byte accesses, ground-truth boxes staged in shared
// memory, one pass over the data, and a grid large enough to cover 148 SMs.
// They are NOT reshaped into GEMMs.
//
// Bit-exactness: every float op below is a single IEEE-rounded operation in
// the same order as the reference's chain of TensorFlow ops (one op per
// tensor expression), so argmax indices / positive masks match the oracle
// exactly.  Built with -fmad=false; critical expressions also use the
// __f*_rn intrinsics, which the compiler never contracts.

#include "common.cuh"

namespace ssd {

// ------------------------------------------------------------------ priors --
// utils/bbox_utils.py:131-214
struct PriorSpec {
    int   n_maps;
    int   fm[SSD_MAX_FEATURE_MAPS];
    int   n_ar[SSD_MAX_FEATURE_MAPS];
    int   offset[SSD_MAX_FEATURE_MAPS + 1];
    float ar[SSD_MAX_FEATURE_MAPS][SSD_MAX_ASPECT_RATIOS];
};

// python: scale_min + ((scale_max - scale_min) / (m - 1)) * (k - 1), float64
__device__ __forceinline__ double prior_scale(int k, int m) {
    const double smin = 0.2, smax = 0.9;
    double step = __ddiv_rn(__dsub_rn(smax, smin), (double)(m - 1));
    return __dadd_rn(smin, __dmul_rn(step, (double)(k - 1)));
}

__global__ void prior_boxes_kernel(PriorSpec s, float4* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int m = 0;
    while (m + 1 < s.n_maps && i >= s.offset[m + 1]) ++m;
    int local = i - s.offset[m];
    int A = s.n_ar[m] + 1;
    int cell = local / A, a = local - cell * A;
    int fm = s.fm[m];
    int y = cell / fm, x = cell - y * fm;                    // y-major cells (:202-205)

    double s_cur = prior_scale(m + 1, s.n_maps);
    float h, w;
    if (a < s.n_ar[m]) {                                     // :169-172 (float32 sqrt/div/mul)
        float root = __fsqrt_rn(s.ar[m][a]);
        h = fdiv(__double2float_rn(s_cur), root);
        w = fmul(__double2float_rn(s_cur), root);
    } else {                                                 // :174-175 (float64 product, float32 sqrt)
        double s_next = prior_scale(m + 2, s.n_maps);
        h = w = __fsqrt_rn(__double2float_rn(__dmul_rn(s_cur, s_next)));
    }
    // :197-201  float64 centre, then one rounding to float32
    double stride = __ddiv_rn(1.0, (double)fm);
    double half = __ddiv_rn(stride, 2.0);
    float cy = __double2float_rn(__dadd_rn(__ddiv_rn((double)y, (double)fm), half));
    float cx = __double2float_rn(__dadd_rn(__ddiv_rn((double)x, (double)fm), half));
    float hh = fdiv(h, 2.0f), hw = fdiv(w, 2.0f);
    float4 r;
    r.x = fminf(fmaxf(fadd(-hh, cy), 0.0f), 1.0f);
    r.y = fminf(fmaxf(fadd(-hw, cx), 0.0f), 1.0f);
    r.z = fminf(fmaxf(fadd(hh, cy), 0.0f), 1.0f);
    r.w = fminf(fmaxf(fadd(hw, cx), 0.0f), 1.0f);
    out[i] = r;
}

// --------------------------------------------------------------------- IoU --
// One IoU exactly as utils/bbox_utils.py:43-55 evaluates it.
__device__ __forceinline__ float iou_ref(const float4 b, float b_area, const float4 g, float g_area) {
    float x_top = fmaxf(b.y, g.y);
    float y_top = fmaxf(b.x, g.x);
    float x_bot = fminf(b.w, g.w);
    float y_bot = fminf(b.z, g.z);
    float inter = fmul(fmaxf(fsub(x_bot, x_top), 0.0f), fmaxf(fsub(y_bot, y_top), 0.0f));
    float uni = fsub(fadd(b_area, g_area), inter);
    return fdiv(inter, uni);
}
__device__ __forceinline__ float box_area(const float4 b) {     // (y2-y1)*(x2-x1)
    return fmul(fsub(b.z, b.x), fsub(b.w, b.y));
}

// grid: (chunks over N*G/VEC, B).  Each thread produces VEC consecutive
// outputs of one anchor row (VEC divides G), stored with one vector store.
template <int VEC>
__global__ void __launch_bounds__(256)
iou_map_kernel(const float4* __restrict__ boxes, const float4* __restrict__ gt, int N, int G,
               int boxes_batched, float* __restrict__ out) {
    extern __shared__ float4 s_gt[];                 // G boxes, then G areas (as float)
    float* s_area = reinterpret_cast<float*>(s_gt + G);
    const int b = blockIdx.y;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float4 v = gt[(size_t)b * G + g];
        s_gt[g] = v;
        s_area[g] = box_area(v);
    }
    __syncthreads();
    const int gv = G / VEC;                          // vectors per anchor row
    const int64_t total = (int64_t)N * gv;
    const float4* bx = boxes + (boxes_batched ? (size_t)b * N : 0);
    float* o = out + (size_t)b * N * G;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(e / gv);
        int g0 = (int)(e - (int64_t)n * gv) * VEC;
        float4 p = __ldg(bx + n);
        float pa = box_area(p);
        float r[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) r[k] = iou_ref(p, pa, s_gt[g0 + k], s_area[g0 + k]);
        float* dst = o + (size_t)n * G + g0;
        if (VEC == 4)      __stcs(reinterpret_cast<float4*>(dst), make_float4(r[0], r[1], r[2], r[3]));
        else if (VEC == 2) __stcs(reinterpret_cast<float2*>(dst), make_float2(r[0], r[VEC > 1 ? 1 : 0]));
        else               __stcs(dst, r[0]);
    }
}

// Warp-tile variant for G <= kIouTileMaxG: a warp owns 32 consecutive anchors and
// all G ground-truth boxes.  Ground-truth boxes that cannot overlap ANY of the
// warp's anchors (tested against the warp's bounding box, one ballot per 32
// boxes) are known to give IoU = +0 and skip the arithmetic entirely, so the
// kernel does ~26 instructions only for (warp, box) pairs that can intersect
// and is otherwise a pure streaming write.  The [32][G] tile is transposed
// through shared memory so the global stores are contiguous 16-byte vectors.
constexpr int kIouTileMaxG = 128;
constexpr int kIouTileWarps = 4;

// warp-wide float min/max in ONE instruction: sm_100a has redux.sync on f32 (CREDUX.MIN/MAX.F32);
// NaN inputs are ignored like fminf/fmaxf.
__device__ __forceinline__ float warp_min_f(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_max_f(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// true when ground-truth box q provably has zero intersection with every box inside
// the bounding box (y1,x1,y2,x2) AND the reference formula then yields exactly +0
// (needs union > 0: anchors with positive area, ground truth with non-negative area).
// Any NaN makes the test false, i.e. the pair is computed exactly.
__device__ __forceinline__ bool gt_culled(const float4 q, float q_area, float wy1, float wx1, float wy2, float wx2) {
    return (q.z <= wy1 || q.x >= wy2 || q.w <= wx1 || q.y >= wx2) && q_area >= 0.0f;
}

// Shared-memory tile of a warp: the [rows][G] block of IoUs in the SAME flat order as global
// memory (so the write-out is a straight 16-byte copy), shifted by `mis` floats so that shared
// and global 16-byte groups coincide, and XOR-swizzled at 16-byte granularity so that the
// column writes of the compute phase (lane stride = G floats) spread over the banks.
__device__ __forceinline__ int tile_swz(int i) { return i ^ (((i >> 5) & 7) << 2); }

template <int NCH>                                   // NCH = ceil(G / 32) chunks of ground-truth boxes
__global__ void __launch_bounds__(kIouTileWarps * 32)
iou_map_tile_kernel(const float4* __restrict__ boxes, const float4* __restrict__ gt, int N, int G, int tile_floats,
                    int boxes_batched, float* __restrict__ out) {
    extern __shared__ float4 s_gt[];                 // [G] boxes | [G] areas (padded to 16 B) | per-warp tiles
    float* s_area = reinterpret_cast<float*>(s_gt + G);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = s_area + ((G + 3) & ~3) + (size_t)warp * tile_floats;
    const int b = blockIdx.y;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float4 v = gt[(size_t)b * G + g];
        s_gt[g] = v;
        s_area[g] = box_area(v);
    }
    __syncthreads();
    const float4* bx = boxes + (boxes_batched ? (size_t)b * N : 0);
    float* o = out + (size_t)b * N * G;
    const int ntiles = (N + 31) >> 5;
    const float inf = __int_as_float(0x7f800000);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = blockIdx.x * kIouTileWarps + warp; t < ntiles; t += gridDim.x * kIouTileWarps) {
        const int n0 = t << 5, n = n0 + lane;
        const bool valid = n < N;
        const float4 p = valid ? __ldg(bx + n) : make_float4(inf, inf, -inf, -inf);
        const float pa = box_area(p);
        const float wy1 = warp_min_f(p.x), wx1 = warp_min_f(p.y), wy2 = warp_max_f(p.z), wx2 = warp_max_f(p.w);
        const bool can_cull = __all_sync(0xffffffffu, !valid || pa > 0.0f);
        uint32_t hitmask[NCH];                           // warp-uniform: which boxes must be evaluated
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int g = (k << 5) + lane;
            bool hit = g < G;
            if (can_cull && hit) hit = !gt_culled(s_gt[g], s_area[g], wy1, wx1, wy2, wx2);
            hitmask[k] = __ballot_sync(0xffffffffu, hit);
            any |= hitmask[k];
        }
        // The tile's outputs are ONE contiguous run of total = rows*G floats.  e = element index,
        // i = e + mis its position relative to the preceding 16-byte boundary of global memory.
        float* dst = o + (size_t)n0 * G;
        const int total = min(32, N - n0) * G;
        const int mis = (int)(((uintptr_t)dst >> 2) & 3);
        const int head = min(total, (4 - mis) & 3);          // scalar elements before the first full group
        const int nvec = (total - head) >> 2;
        const int tail0 = head + (nvec << 2);
        float4* vdst = reinterpret_cast<float4*>(dst + head);
        if (any == 0) {                                      // nothing can overlap: stream +0
            if (lane < head) __stcs(dst + lane, 0.0f);
            for (int v = lane; v < nvec; v += 32) __stcs(vdst + v, zero4);
            if (tail0 + lane < total) __stcs(dst + tail0 + lane, 0.0f);
            continue;
        }
        float4* t4 = reinterpret_cast<float4*>(tile);
        for (int v = lane; v < (tile_floats >> 2); v += 32) t4[v] = zero4;
        __syncwarp();
        const int rowbase = lane * G + mis;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            uint32_t m = hitmask[k];
            while (m) {
                const int gg = (k << 5) + __ffs(m) - 1;
                m &= m - 1;
                const float v = iou_ref(p, pa, s_gt[gg], s_area[gg]);
                tile[tile_swz(rowbase + gg)] = v;            // rows of lanes past N are never written out
            }
        }
        __syncwarp();
        if (lane < head) __stcs(dst + lane, tile[tile_swz(mis + lane)]);
        {
            const int g0 = (mis + head) >> 2;                // first full 16-byte group of the tile
            for (int v = lane; v < nvec; v += 32)
                __stcs(vdst + v, t4[tile_swz((g0 + v) << 2) >> 2]);
        }
        if (tail0 + lane < total) __stcs(dst + tail0 + lane, tile[tile_swz(mis + tail0 + lane)]);
        __syncwarp();
    }
}

// ---------------------------------------------------------- encode / decode --
// utils/bbox_utils.py:95-128 for one (prior, matched gt) pair -> [dy,dx,dh,dw]
__device__ __forceinline__ float4 encode_one(const float4 p, const float4 g) {
    float pw = fsub(p.w, p.y), ph = fsub(p.z, p.x);
    float pcx = fadd(p.y, fmul(0.5f, pw)), pcy = fadd(p.x, fmul(0.5f, ph));
    float gw = fsub(g.w, g.y), gh = fsub(g.z, g.x);
    float gcx = fadd(g.y, fmul(0.5f, gw)), gcy = fadd(g.x, fmul(0.5f, gh));
    if (pw == 0.0f) pw = 1e-3f;
    if (ph == 0.0f) ph = 1e-3f;
    float4 d;
    d.x = (gh == 0.0f) ? 0.0f : fdiv(fsub(gcy, pcy), ph);
    d.y = (gw == 0.0f) ? 0.0f : fdiv(fsub(gcx, pcx), pw);
    d.z = (gh == 0.0f) ? 0.0f : logf(fdiv(gh, ph));
    d.w = (gw == 0.0f) ? 0.0f : logf(fdiv(gw, pw));
    return d;
}

// utils/bbox_utils.py:68-82 for one (prior, delta) pair -> [y1,x1,y2,x2]
__device__ __forceinline__ float4 decode_one(const float4 p, const float4 d) {
    float pw = fsub(p.w, p.y), ph = fsub(p.z, p.x);
    float pcx = fadd(p.y, fmul(0.5f, pw)), pcy = fadd(p.x, fmul(0.5f, ph));
    float w = fmul(expf(d.w), pw), h = fmul(expf(d.z), ph);
    float cx = fadd(fmul(d.y, pw), pcx), cy = fadd(fmul(d.x, ph), pcy);
    float y1 = fsub(cy, fmul(0.5f, h)), x1 = fsub(cx, fmul(0.5f, w));
    return make_float4(y1, x1, fadd(h, y1), fadd(w, x1));
}

__global__ void __launch_bounds__(256)
encode_kernel(const float4* __restrict__ priors, const float4* __restrict__ boxes, int64_t total, int N,
              int priors_batched, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = __ldg(priors + (priors_batched ? i : i % N));
        out[i] = encode_one(p, __ldcs(boxes + i));
    }
}
__global__ void __launch_bounds__(256)
decode_kernel(const float4* __restrict__ priors, const float4* __restrict__ deltas, int64_t total, int N,
              int priors_batched, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = __ldg(priors + (priors_batched ? i : i % N));
        out[i] = decode_one(p, __ldcs(deltas + i));
    }
}

// ------------------------------------------------------------ match+encode --
// utils/train_utils.py:123-135 fused with the encode above.  One thread per
// (image, anchor); the G ground-truth boxes of the image live in shared
// memory; the IoU row exists only in registers.
//
// Culling (warp granularity, one ballot per 32 ground-truth boxes):
//  * a box disjoint from the warp's bounding box has IoU = +0 for all 32 anchors and can
//    never win the strict ">" of the arg-max;
//  * NEED_IDX = false (the arg-max index itself is not an output): only anchors whose best
//    IoU exceeds iou_thr produce anything that depends on the index, so a box that provably
//    cannot reach iou_thr with ANY anchor of the warp is skipped as well.  IoU > t implies
//    inter > t*area_gt, inter > t*area_anchor, and inter <= area(gt ^ warp bounding box),
//    inter <= min(area_gt, area_anchor); the tests below use a 10 % margin, far above the
//    few-ulp rounding of the float32 expressions, so no box with IoU > t is ever skipped and
//    the surviving boxes are evaluated with the exact reference arithmetic in index order.
//
// The one-hot block of the CTA (kMatchThreads * L consecutive floats) is built in shared
// memory (zero fill, one store per anchor) and streamed out with 16-byte stores: the
// dominant 4*L bytes/anchor of this kernel.
constexpr int kMatchThreads = 256;

template <bool NEED_IDX, bool SMEM_ONEHOT>
__global__ void __launch_bounds__(kMatchThreads)         // (forcing 32 registers for full occupancy measured 3-12 % slower)
match_encode_kernel(const float4* __restrict__ priors, const float4* __restrict__ gt_boxes,
                    const int32_t* __restrict__ gt_labels, int N, int G, int L, float iou_thr,
                    float4 variances, float4* __restrict__ out_deltas, float* __restrict__ out_onehot,
                    int32_t* __restrict__ out_label, int32_t* __restrict__ out_match) {
    extern __shared__ float4 s_gt[];                               // [G] boxes
    float*   s_area = reinterpret_cast<float*>(s_gt + G);          // [G]
    int32_t* s_glab = reinterpret_cast<int32_t*>(s_area + G);      // [G]
    int32_t* s_lab  = s_glab + G;                                  // [kMatchThreads]
    // [kMatchThreads*L + 4] one-hot staging, 16-byte aligned (SMEM_ONEHOT only)
    float*   s_hot  = reinterpret_cast<float*>(s_gt) + ((6 * G + kMatchThreads + 3) & ~3);
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kMatchThreads;
    const int cnt = min(kMatchThreads, N - n0);
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float4 v = gt_boxes[(size_t)b * G + g];
        s_gt[g] = v;
        s_area[g] = box_area(v);
        s_glab[g] = gt_labels[(size_t)b * G + g];
    }
    if (SMEM_ONEHOT && out_onehot != nullptr) {
        float4* h4 = reinterpret_cast<float4*>(s_hot);
        const int n4 = (cnt * L + 3 + 3) >> 2;
        for (int v = threadIdx.x; v < n4; v += kMatchThreads) h4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    const int n = n0 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = n < N;
    const float inf = __int_as_float(0x7f800000);
    int label = 0;
    {
        const float4 p = valid ? __ldg(priors + n) : make_float4(inf, inf, -inf, -inf);
        const float pa = box_area(p);
        const float wy1 = warp_min_f(p.x), wx1 = warp_min_f(p.y), wy2 = warp_max_f(p.z), wx2 = warp_max_f(p.w);
        const bool can_cull = __all_sync(0xffffffffu, !valid || pa > 0.0f);
        float best;
        int idx = 0;
        if (can_cull) {
            best = 0.0f;                                           // == IoU of every disjoint box
            // threshold culling needs 0 < thr (and finite areas); a_lo/a_hi bound the warp's anchor areas
            const bool thr_cull = !NEED_IDX && iou_thr > 0.0f;
            const float a_lo = warp_min_f(valid ? pa : inf), a_hi = warp_max_f(valid ? pa : 0.0f);
            const float t9 = 0.9f * iou_thr;
            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                bool hit = false;
                if (g < G) {
                    const float4 q = s_gt[g];
                    const float qa = s_area[g];
                    hit = !gt_culled(q, qa, wy1, wx1, wy2, wx2);
                    if (thr_cull && hit) {
                        const float cy = fminf(q.z, wy2) - fmaxf(q.x, wy1), cx = fminf(q.w, wx2) - fmaxf(q.y, wx1);
                        const float cover = cy * cx;               // area(gt ^ warp bounding box) >= any inter
                        if (cover < t9 * qa || cover < t9 * a_lo || a_hi < t9 * qa || qa < t9 * a_lo) hit = false;
                    }
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int gg = g0 + __ffs(mask) - 1;
                    mask &= mask - 1;
                    // iou_ref's intersection / union, with the division only when the quotient can exceed `best`:
                    // inter / uni > best needs inter > best * uni; the test keeps a 2^-17 relative margin (far above
                    // the one-ulp roundings of the product), so every candidate that could win the strict ">" is
                    // still divided exactly like the reference does and the arg-max is unchanged.  (uni <= 0 or NaN
                    // -- inverted ground-truth boxes -- fail or pass the test consistently with v > best: 0 > 0 is
                    // false either way, a negative bound sends the pair to the exact path.)
                    const float4 q = s_gt[gg];
                    const float inter = fmul(fmaxf(fsub(fminf(p.w, q.w), fmaxf(p.y, q.y)), 0.0f),
                                             fmaxf(fsub(fminf(p.z, q.z), fmaxf(p.x, q.x)), 0.0f));
                    const float uni = fsub(fadd(pa, s_area[gg]), inter);
                    if (inter > fmul(fmul(best, uni), 0.99999f)) {
                        const float v = fdiv(inter, uni);
                        if (v > best) { best = v; idx = gg; }      // ascending g: first maximum wins (:124)
                    }
                }
            }
        } else {                                                   // degenerate anchors: plain scan (0/0 = NaN semantics)
            best = iou_ref(p, pa, s_gt[0], s_area[0]);
            for (int g = 1; g < G; ++g) {
                float v = iou_ref(p, pa, s_gt[g], s_area[g]);
                if (v > best) { best = v; idx = g; }
            }
        }
        const bool pos = valid && best > iou_thr;                  // strict (:126)
        // encode(p, 0-box) / variances is exactly +0 for positive finite variances: only warps
        // that hold a positive anchor run the log/divide chain
        const bool var_ok = variances.x > 0.f && variances.y > 0.f && variances.z > 0.f && variances.w > 0.f;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (__any_sync(0xffffffffu, pos) || !var_ok) {
            float4 g4 = pos ? s_gt[idx] : make_float4(0.f, 0.f, 0.f, 0.f);   // :129-130
            d = encode_one(p, g4);
            d.x = fdiv(d.x, variances.x); d.y = fdiv(d.y, variances.y);      // :131
            d.z = fdiv(d.z, variances.z); d.w = fdiv(d.w, variances.w);
        }
        if (valid) {
            __stcs(out_deltas + (size_t)b * N + n, d);
            label = pos ? s_glab[idx] : 0;                             // :133-134
            if (out_label) out_label[(size_t)b * N + n] = label;
            if (NEED_IDX && out_match) out_match[(size_t)b * N + n] = idx;
        }
    }
    if (out_onehot == nullptr) return;

    // :135  one_hot: CTA writes floats [first, first + cnt*L) of the output.
    const int64_t first = ((int64_t)b * N + n0) * L;
    const int total = cnt * L;
    float* dst = out_onehot + first;
    if (SMEM_ONEHOT) {
        float* hot = stage_rows_ptr(dst, s_hot);                   // same 16-byte phase as global memory
        if (valid && label >= 0 && label < L) hot[threadIdx.x * L + label] = 1.0f;
        __syncthreads();
        stage_rows_out(dst, total, hot);
        return;
    }
    s_lab[threadIdx.x] = label;
    __syncthreads();
    const int head = (int)((4 - (((uintptr_t)dst >> 2) & 3)) & 3);      // floats until 16B alignment
    const int nvec = (total > head) ? (total - head) >> 2 : 0;
    for (int e = threadIdx.x; e < min(head, total); e += blockDim.x) {
        int a = e / L;
        dst[e] = (s_lab[a] == e - a * L) ? 1.0f : 0.0f;
    }
    {
        // (anchor, label) of the thread's first element by one division, then advanced
        // incrementally by the block stride (4 * blockDim.x elements per step)
        const int step = kMatchThreads * 4, step_a = step / L, step_l = step - step_a * L;
        int e = head + (threadIdx.x << 2);
        int a = e / L, l = e - a * L;
        for (int v = threadIdx.x; v < nvec; v += kMatchThreads) {
            float r[4];
            int aa = a, ll = l;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = (s_lab[aa] == ll) ? 1.0f : 0.0f;
                if (++ll == L) { ll = 0; ++aa; }
            }
            __stcs(reinterpret_cast<float4*>(dst + e), make_float4(r[0], r[1], r[2], r[3]));
            e += step; a += step_a; l += step_l;
            if (l >= L) { l -= L; ++a; }
        }
    }
    for (int e = head + (nvec << 2) + threadIdx.x; e < total; e += blockDim.x) {
        int a = e / L;
        dst[e] = (s_lab[a] == e - a * L) ? 1.0f : 0.0f;
    }
}

}  // namespace ssd

// =============================================================== C entries ==
using namespace ssd;

extern "C" int ssd_prior_box_count(const int* h_fm_shapes, int n_maps, const int* h_ar_counts) {
    if (!h_fm_shapes || !h_ar_counts || n_maps < 1 || n_maps > SSD_MAX_FEATURE_MAPS) return -1;
    long total = 0;
    for (int i = 0; i < n_maps; ++i) {
        if (h_fm_shapes[i] < 1 || h_ar_counts[i] < 0 || h_ar_counts[i] > SSD_MAX_ASPECT_RATIOS) return -1;
        total += (long)h_fm_shapes[i] * h_fm_shapes[i] * (h_ar_counts[i] + 1);
    }
    return total > 0x7fffffffL ? -1 : (int)total;
}

extern "C" int ssd_prior_boxes(const int* h_fm_shapes, int n_maps, const float* h_aspect_ratios,
                               const int* h_ar_counts, float* d_out, int n_anchors, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(h_fm_shapes); SSD_REQUIRE_PTR(h_aspect_ratios); SSD_REQUIRE_PTR(h_ar_counts);
    SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(n_maps >= 2 && n_maps <= SSD_MAX_FEATURE_MAPS, SSD_ERR_SHAPE,
                "ssd_prior_boxes: n_maps=%d outside [2,%d]", n_maps, SSD_MAX_FEATURE_MAPS);
    int expect = ssd_prior_box_count(h_fm_shapes, n_maps, h_ar_counts);
    SSD_REQUIRE(expect > 0 && expect == n_anchors, SSD_ERR_SHAPE,
                "ssd_prior_boxes: n_anchors=%d but the feature maps define %d", n_anchors, expect);
    PriorSpec s{};
    s.n_maps = n_maps;
    int off = 0, k = 0;
    for (int i = 0; i < n_maps; ++i) {
        s.fm[i] = h_fm_shapes[i];
        s.n_ar[i] = h_ar_counts[i];
        s.offset[i] = off;
        for (int a = 0; a < h_ar_counts[i]; ++a) s.ar[i][a] = h_aspect_ratios[k++];
        off += h_fm_shapes[i] * h_fm_shapes[i] * (h_ar_counts[i] + 1);
    }
    s.offset[n_maps] = off;
    prior_boxes_kernel<<<ceil_div(n_anchors, 256), 256, 0, as_stream(stream)>>>(
        s, reinterpret_cast<float4*>(d_out), n_anchors);
    SSD_CHECK_LAUNCH("prior_boxes_kernel");
    return SSD_OK;
}

extern "C" int ssd_iou_map(const float* d_boxes, const float* d_gt, int B, int N, int G, int boxes_batched,
                           float* d_out, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_boxes); SSD_REQUIRE_PTR(d_gt); SSD_REQUIRE_PTR(d_out);
    SSD_REQUIRE(B >= 0 && N >= 0 && G >= 0 && B <= 65535, SSD_ERR_SHAPE,
                "ssd_iou_map: bad shape B=%d N=%d G=%d", B, N, G);
    if (B == 0 || N == 0 || G == 0) return SSD_OK;
    if (G <= kIouTileMaxG) {
        // per-warp tile: 32*G floats + up to 3 of misalignment, rounded to the swizzle block (32 floats)
        const int tile_floats = (32 * G + 3 + 31) & ~31;
        size_t smem_t = (size_t)G * 16 + (size_t)((G + 3) & ~3) * 4 + (size_t)kIouTileWarps * tile_floats * sizeof(float);
        const int nch = (G + 31) / 32;
        auto kern = nch == 1 ? iou_map_tile_kernel<1> : nch == 2 ? iou_map_tile_kernel<2>
                  : nch == 3 ? iou_map_tile_kernel<3> : iou_map_tile_kernel<4>;
        if (smem_t > 40 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
        const int ntiles = (N + 31) / 32;
        const int per_img = (ntiles + kIouTileWarps - 1) / kIouTileWarps;
        const int cap = max(1, (sm_count() * 32 + B - 1) / B);
        dim3 grid(min(per_img, cap), B);
        kern<<<grid, kIouTileWarps * 32, smem_t, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(d_boxes), reinterpret_cast<const float4*>(d_gt), N, G, tile_floats,
            boxes_batched, d_out);
        SSD_CHECK_LAUNCH("iou_map_tile_kernel");
        return SSD_OK;
    }
    size_t smem = (size_t)G * 20;
    SSD_REQUIRE(smem <= 160 * 1024, SSD_ERR_UNSUPPORTED, "ssd_iou_map: G=%d exceeds shared-memory staging", G);
    const bool aligned = (((uintptr_t)d_out) & 15) == 0;
    int vec = (aligned && G % 4 == 0) ? 4 : ((((uintptr_t)d_out) & 7) == 0 && G % 2 == 0) ? 2 : 1;
    int64_t work = (int64_t)N * (G / vec);
    int per_img = (int)((work + 255) / 256);
    // enough CTAs to cover 148 SMs x 8 resident CTAs across the B images, no more
    int cap = max(1, (sm_count() * 8 + B - 1) / B);
    dim3 grid(min(per_img, cap), B);
    auto launch = [&](auto kern) {
        if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, 256, smem, as_stream(stream)>>>(reinterpret_cast<const float4*>(d_boxes),
                                                     reinterpret_cast<const float4*>(d_gt), N, G,
                                                     boxes_batched, d_out);
    };
    if (vec == 4) launch(iou_map_kernel<4>);
    else if (vec == 2) launch(iou_map_kernel<2>);
    else launch(iou_map_kernel<1>);
    SSD_CHECK_LAUNCH("iou_map_kernel");
    return SSD_OK;
}

extern "C" int ssd_match_encode(const float* d_priors, const float* d_gt_boxes, const int32_t* d_gt_labels,
                                int B, int N, int G, int L, float iou_threshold, const float* h_variances,
                                float* d_deltas, float* d_onehot, int32_t* d_label, int32_t* d_match,
                                ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_priors); SSD_REQUIRE_PTR(d_gt_boxes); SSD_REQUIRE_PTR(d_gt_labels);
    SSD_REQUIRE_PTR(h_variances); SSD_REQUIRE_PTR(d_deltas);
    SSD_REQUIRE(B >= 0 && N >= 0 && G >= 1 && L >= 1 && B <= 65535, SSD_ERR_SHAPE,
                "ssd_match_encode: bad shape B=%d N=%d G=%d L=%d (G must be >= 1)", B, N, G, L);
    if (B == 0 || N == 0) return SSD_OK;
    const size_t smem_base = (size_t)((6 * G + kMatchThreads + 3) & ~3) * 4;      // gt boxes/areas/labels + anchor labels
    const size_t smem_hot = ((size_t)kMatchThreads * L + 8) * sizeof(float);        // one-hot staging
    SSD_REQUIRE(smem_base <= 160 * 1024, SSD_ERR_UNSUPPORTED, "ssd_match_encode: G=%d exceeds shared-memory staging", G);
    const bool smem_onehot = d_onehot != nullptr && smem_base + smem_hot <= 100 * 1024;   // >= 2 CTAs per SM
    const size_t smem = smem_base + (smem_onehot ? smem_hot : 0);
    dim3 grid(ceil_div(N, kMatchThreads), B);
    float4 var = make_float4(h_variances[0], h_variances[1], h_variances[2], h_variances[3]);
    auto launch = [&](auto kern) {
        if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kMatchThreads, smem, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(d_priors), reinterpret_cast<const float4*>(d_gt_boxes), d_gt_labels,
            N, G, L, iou_threshold, var, reinterpret_cast<float4*>(d_deltas), d_onehot, d_label, d_match);
    };
    if (d_match != nullptr) { if (smem_onehot) launch(match_encode_kernel<true, true>); else launch(match_encode_kernel<true, false>); }
    else                    { if (smem_onehot) launch(match_encode_kernel<false, true>); else launch(match_encode_kernel<false, false>); }
    SSD_CHECK_LAUNCH("match_encode_kernel");
    return SSD_OK;
}

static int elementwise_grid(int64_t total) {
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < cap ? blocks : cap);
}

extern "C" int ssd_encode_deltas(const float* d_priors, const float* d_boxes, int B, int N, int priors_batched,
                                 float* d_deltas, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_priors); SSD_REQUIRE_PTR(d_boxes); SSD_REQUIRE_PTR(d_deltas);
    SSD_REQUIRE(B >= 0 && N >= 0, SSD_ERR_SHAPE, "ssd_encode_deltas: bad shape B=%d N=%d", B, N);
    int64_t total = (int64_t)B * N;
    if (total == 0) return SSD_OK;
    encode_kernel<<<elementwise_grid(total), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(d_priors), reinterpret_cast<const float4*>(d_boxes), total, N,
        priors_batched, reinterpret_cast<float4*>(d_deltas));
    SSD_CHECK_LAUNCH("encode_kernel");
    return SSD_OK;
}

extern "C" int ssd_decode_boxes(const float* d_priors, const float* d_deltas, int B, int N, int priors_batched,
                                float* d_boxes, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_priors); SSD_REQUIRE_PTR(d_deltas); SSD_REQUIRE_PTR(d_boxes);
    SSD_REQUIRE(B >= 0 && N >= 0, SSD_ERR_SHAPE, "ssd_decode_boxes: bad shape B=%d N=%d", B, N);
    int64_t total = (int64_t)B * N;
    if (total == 0) return SSD_OK;
    decode_kernel<<<elementwise_grid(total), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(d_priors), reinterpret_cast<const float4*>(d_deltas), total, N,
        priors_batched, reinterpret_cast<float4*>(d_boxes));
    SSD_CHECK_LAUNCH("decode_kernel");
    return SSD_OK;
}
