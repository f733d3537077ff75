// Decode + combined NMS (models/decoder.py:60-93 -> utils/bbox_utils.py:10-21
// -> tf.image.combined_non_max_suppression) and softmax (models/header.py:88)
// for sm_100a.  Everything stays on the device: the reference's NMS is a
// CPU-only TensorFlow kernel that forces a device->host copy of all head
// outputs.
//
// Two kernels per call:
//   A. candidate pass -- full grid, one thread per (image, anchor); score rows
//      staged through shared memory; optional fused softmax; background-row
//      rule; survivors (score > threshold) appended to a per-image list as
//      64-bit keys  [class:8 | ~order(score):32 | anchor:24].
//   B. per-image pass -- one CTA per image: bitonic sort of the keys (class
//      asc, score desc, anchor asc), greedy per-class suppression with one
//      warp per class, then a second sort of the survivors by (score desc,
//      class asc, anchor asc) and emission of the first max_total rows.
// Boxes are never materialised for all anchors: pass B decodes the (few)
// candidate boxes on demand from priors+deltas with the same arithmetic.
//
// Tie order on equal scores follows the oracle rule documented in
// oracle/box_oracle.py (TensorFlow leaves it implementation-defined).

#include "common.cuh"

namespace ssd {

constexpr int kRowThreadsNms = 128;
constexpr int kNmsThreads = 512;          // 64 registers x 512 threads: two images per SM (1024 threads filled the register file)
constexpr uint64_t kPadKey = ~0ull;

__device__ __forceinline__ uint32_t order_bits(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

__device__ __forceinline__ float exp2f_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float warp_max_f32(float v) {     // redux.sync on f32: sm_100a
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// ---------------------------------------------------------------- softmax --
__global__ void __launch_bounds__(kRowThreadsNms)
softmax_kernel(const float* __restrict__ logits, int64_t rows, int L, float* __restrict__ probs) {
    extern __shared__ __align__(16) float s_rows[];
    for (int64_t r0 = (int64_t)blockIdx.x * kRowThreadsNms; r0 < rows; r0 += (int64_t)gridDim.x * kRowThreadsNms) {
        const int cnt = (int)min((int64_t)kRowThreadsNms, rows - r0);
        float* rows = stage_rows_in(logits + r0 * L, cnt * L, s_rows);   // probs + r0*L has the same misalignment
        stage_rows_wait();
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            float* z = rows + (size_t)threadIdx.x * L;
            float m = z[0];
            for (int l = 1; l < L; ++l) m = fmaxf(m, z[l]);
            float s = 0.0f;
            for (int l = 0; l < L; ++l) { float e = expf(fsub(z[l], m)); z[l] = e; s = fadd(s, e); }
            for (int l = 0; l < L; ++l) z[l] = fdiv(z[l], s);
        }
        __syncthreads();
        float* dst = probs + r0 * L;
        if (stage_rows_ptr(dst, s_rows) == rows) {
            stage_rows_out(dst, cnt * L, rows);
        } else {                                          // input and output misaligned differently: scalar stores
            for (int e = threadIdx.x; e < cnt * L; e += blockDim.x) __stcs(dst + e, rows[e]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------- candidate pass --
// DECODER=true : models/decoder.py:78-83 rule (argmax==0 kills the row).
// DECODER=false: plain combined NMS, every (anchor, class) above threshold.
//
// FROM_LOGITS (models/header.py:88 fused in) runs in two phases so that the full-precision
// softmax is only evaluated for rows that can produce a candidate:
//   1. per thread, one row: max, then an approximate sum of exponentials (ex2.approx, relative
//      error < 1e-5).  The largest probability of the row is 1/sum, so a row with
//      1/sum < thr * (1 - 1e-4) certainly has no score above thr; with DECODER a row whose
//      first maximum is class 0 is certainly zeroed by the background rule.
//   2. per warp, the surviving rows one at a time with the classes spread over the lanes: the
//      arithmetic of softmax_kernel (expf, sequential float32 sum, IEEE division), so the
//      fused path emits bit-identical scores to softmax followed by the probability path.
template <bool DECODER, bool FROM_LOGITS, int LT>          // LT: compile-time label count (0 = use the runtime L)
__global__ void __launch_bounds__(kRowThreadsNms)
nms_candidates_kernel(const float* __restrict__ scores, int N, int L_rt, float score_thr, int cap,
                      uint64_t* __restrict__ keys, int key_stride, int* __restrict__ counts) {
    const int L = LT ? LT : L_rt;                        // phase 1 unrolls completely for the VOC label count
    extern __shared__ __align__(16) float s_rows[];
    __shared__ int s_warp_tot[kRowThreadsNms / 32];
    __shared__ int s_base;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kRowThreadsNms;
    const int cnt = min(kRowThreadsNms, N - n0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* rows = stage_rows_in(scores + ((size_t)b * N + n0) * L, cnt * L, s_rows);
    stage_rows_wait();
    __syncthreads();
    float* p = rows + (size_t)threadIdx.x * L;
    int mine = 0;                                        // candidates of this anchor
    int first_cls = -1;                                  // its lowest candidate class when known (fused-softmax path)
    if (FROM_LOGITS) {
        bool maybe = false;
        if ((int)threadIdx.x < cnt) {
            float m = p[0];
#pragma unroll
            for (int l = 1; l < L; ++l) m = fmaxf(m, p[l]);
            float s = 0.0f;
#pragma unroll
            for (int l = 0; l < L; ++l) s += exp2f_approx((p[l] - m) * 1.4426950408889634f);
            const float thr_lo = score_thr - 1e-4f * fabsf(score_thr);
            maybe = thr_lo * s < 1.0f;
            if (DECODER && p[0] == m) maybe = false;     // first maximum is the background class
            if (!(s == s)) maybe = true;                 // NaN/Inf rows: let the exact path decide
        }
        unsigned todo = __ballot_sync(0xffffffffu, maybe);
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            float* q = rows + (size_t)((wid << 5) + r) * L;
            float m = -__int_as_float(0x7f800000);
            for (int l = lane; l < L; l += 32) m = fmaxf(m, q[l]);
            m = warp_max_f32(m);
            for (int l = lane; l < L; l += 32) q[l] = expf(fsub(q[l], m));
            __syncwarp();
            float s = 0.0f;
            for (int l = 0; l < L; ++l) s = fadd(s, q[l]);           // same order as softmax_kernel
            __syncwarp();
            int c = 0, am = 0x7fffffff, lowest = 0x7fffffff;
            float best = -__int_as_float(0x7f800000);
            for (int l = lane; l < L; l += 32) {
                const float v = fdiv(q[l], s);
                q[l] = v;
                if (v > score_thr) { ++c; lowest = min(lowest, l); }
                if (v > best) { best = v; am = l; }
            }
            c = __reduce_add_sync(0xffffffffu, c);
            lowest = __reduce_min_sync(0xffffffffu, lowest);
            if (DECODER) {                               // decoder.py:78-83: first maximum == 0 zeroes the row
                const float gb = warp_max_f32(best);
                const int first = __reduce_min_sync(0xffffffffu, best == gb ? am : 0x7fffffff);
                if (first == 0 || first == 0x7fffffff) c = 0;
            }
            if (lane == r) { mine = c; first_cls = lowest; }
            __syncwarp();
        }
    } else if ((int)threadIdx.x < cnt) {
        bool alive = true;
        if (DECODER) {                                   // decoder.py:78-83: argmax == 0 zeroes the row
            int am = 0;
            float best = p[0];
            for (int l = 1; l < L; ++l)
                if (p[l] > best) { best = p[l]; am = l; }
            alive = am != 0;
        }
        if (alive)
            for (int l = 0; l < L; ++l) mine += (p[l] > score_thr) ? 1 : 0;      // strict, like TensorFlow
    }
    // One global reservation per CTA: exclusive scan of the per-anchor counts, a single
    // atomicAdd for the CTA's total (the previous per-candidate atomics serialised on one
    // address per image and stalled every warp for a full L2 round trip each).
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kRowThreadsNms / 32; ++w) { int t = s_warp_tot[w]; s_warp_tot[w] = tot; tot += t; }
        s_base = tot ? atomicAdd(counts + b, tot) : 0;
    }
    __syncthreads();
    if (mine == 0) return;
    int slot = s_base + s_warp_tot[wid] + incl - mine;
    const uint32_t anchor = (uint32_t)(n0 + threadIdx.x);
    uint64_t* dst = keys + (size_t)b * key_stride;
    if (mine == 1 && first_cls >= 0) {                   // the usual case above a 0.5 threshold: exactly one class
        if (slot < cap) dst[slot] = ((uint64_t)first_cls << 56) | ((uint64_t)(~order_bits(p[first_cls])) << 24) | anchor;
        return;
    }
    for (int l = 0; l < L; ++l) {
        float sc = p[l];
        if (sc > score_thr) {
            if (slot < cap) dst[slot] = ((uint64_t)l << 56) | ((uint64_t)(~order_bits(sc)) << 24) | anchor;
            ++slot;
        }
    }
}

// --------------------------------------------------------- per-image pass --
struct DecodeFetch {          // boxes decoded on demand: decoder.py:74-75
    const float4* priors; const float4* deltas; float4 var; int N;
    __device__ __forceinline__ float4 operator()(int b, int anchor, int /*cls*/) const;
};
struct DirectFetch {          // boxes given: [B,N,q,4]
    const float4* boxes; int N; int q;
    __device__ __forceinline__ float4 operator()(int b, int anchor, int cls) const {
        return __ldg(boxes + ((size_t)b * N + anchor) * q + (q > 1 ? cls : 0));
    }
};

// utils/bbox_utils.py:68-82 (same expression order as box_kernels.cu:decode_one)
__device__ __forceinline__ float4 DecodeFetch::operator()(int b, int anchor, int) const {
    float4 p = __ldg(priors + anchor);
    float4 d = __ldg(deltas + (size_t)b * N + anchor);
    d.x = fmul(d.x, var.x); d.y = fmul(d.y, var.y); d.z = fmul(d.z, var.z); d.w = fmul(d.w, var.w);
    float pw = fsub(p.w, p.y), ph = fsub(p.z, p.x);
    float pcx = fadd(p.y, fmul(0.5f, pw)), pcy = fadd(p.x, fmul(0.5f, ph));
    float w = fmul(expf(d.w), pw), h = fmul(expf(d.z), ph);
    float cx = fadd(fmul(d.y, pw), pcx), cy = fadd(fmul(d.x, ph), pcy);
    float y1 = fsub(cy, fmul(0.5f, h)), x1 = fsub(cx, fmul(0.5f, w));
    return make_float4(y1, x1, fadd(h, y1), fadd(w, x1));
}

// [TF-recall] IoU of TensorFlow's non_max_suppression_op.cc: corners canonicalised, 0 if either area <= 0; a candidate
// is suppressed when  inter / (area_a + area_b - inter) > thr.  The test below decides that comparison
// without the division in all but borderline cases: for thr > 0 the quotient inter / uni exceeds thr
// iff inter > thr * uni, and a 1e-5 relative margin (far above the roundings of the product and of the quotient) decides
// every pair that is not within that margin of the threshold; those few take the exact division.
__device__ __forceinline__ bool nms_iou_gt(const float4 a, const float4 b, float thr) {
    float aymin = fminf(a.x, a.z), aymax = fmaxf(a.x, a.z), axmin = fminf(a.y, a.w), axmax = fmaxf(a.y, a.w);
    float bymin = fminf(b.x, b.z), bymax = fmaxf(b.x, b.z), bxmin = fminf(b.y, b.w), bxmax = fmaxf(b.y, b.w);
    float area_a = fmul(fsub(aymax, aymin), fsub(axmax, axmin));
    float area_b = fmul(fsub(bymax, bymin), fsub(bxmax, bxmin));
    if (area_a <= 0.0f || area_b <= 0.0f) return 0.0f > thr;
    float ih = fmaxf(fsub(fminf(aymax, bymax), fmaxf(aymin, bymin)), 0.0f);
    float iw = fmaxf(fsub(fminf(axmax, bxmax), fmaxf(axmin, bxmin)), 0.0f);
    float inter = fmul(ih, iw);
    float uni = fsub(fadd(area_a, area_b), inter);
    if (thr > 0.0f && uni > 0.0f) {
        const float tu = fmul(thr, uni);
        if (inter < fmul(tu, 0.99999f)) return false;
        if (inter > fmul(tu, 1.00001f)) return true;
    }
    return fdiv(inter, uni) > thr;
}

// The same test on boxes whose corners are already canonical (ymin, xmin, ymax, xmax): the shared box cache stores them
// that way, so the per-pair work is the intersection, the union and the comparison only.
__device__ __forceinline__ float4 nms_canonical(const float4 a) {
    return make_float4(fminf(a.x, a.z), fminf(a.y, a.w), fmaxf(a.x, a.z), fmaxf(a.y, a.w));
}
__device__ __forceinline__ bool nms_iou_gt_canonical(const float4 a, const float4 b, float thr) {
    const float area_a = fmul(fsub(a.z, a.x), fsub(a.w, a.y)), area_b = fmul(fsub(b.z, b.x), fsub(b.w, b.y));
    if (area_a <= 0.0f || area_b <= 0.0f) return 0.0f > thr;
    const float ih = fmaxf(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float iw = fmaxf(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = fmul(ih, iw);
    const float uni = fsub(fadd(area_a, area_b), inter);
    if (thr > 0.0f && uni > 0.0f) {
        const float tu = fmul(thr, uni);
        if (inter < fmul(tu, 0.99999f)) return false;
        if (inter > fmul(tu, 1.00001f)) return true;
    }
    return fdiv(inter, uni) > thr;
}

// In-place ascending bitonic sort of a[0..P) (P a power of two) by the CTA.
__device__ void bitonic_sort(uint64_t* a, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t x = a[i], y = a[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}
// Ascending sort of n UNIQUE keys by rank counting: every key counts the smaller keys (broadcast shared-memory
// reads, no barrier per stage) and is written to its rank.  O(n^2) compares but only two barriers: cheaper than
// the 36-55 barrier stages of the bitonic network for the few hundred keys an image produces.
__device__ void rank_sort(uint64_t* a, uint64_t* tmp, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = a[i];
        int rank = 0;
#pragma unroll 8
        for (int j = 0; j < n; ++j) rank += (a[j] < key) ? 1 : 0;
        tmp[rank] = key;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = tmp[i];
    __syncthreads();
}
constexpr int kRankSortMax = 768;          // beyond this the bitonic network wins
constexpr int kMatrixMaxSeg = 256;         // largest class (candidates) handled by the suppression bit matrix

__device__ __forceinline__ int pow2_ceil(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

struct NmsParams {
    int N, L, per_class, max_total, cap;
    int key_stride;            // u64 slots per image in `keys` (power of two >= cap)
    int merge_stride;          // u64 slots per image in `merge` (power of two >= L*per_class)
    int smem_sort_slots;       // u64 slots of the shared sort buffer
    int fast_slots;            // > 0: shared box cache (float4 per candidate) + u16 kept lists are available
    float iou_thr;
    int clip;
    int labels_first;          // output order: decoder (boxes, labels, scores) vs TF (boxes, scores, classes)
    unsigned long long* trace; // debug (ssd_debug_trace): phase boundaries of image 0, else nullptr
};

// One CTA per image.  Fast path (candidate count <= fast_slots): keys, the decoded
// candidate boxes (decoded once, cooperatively, in sorted order) and the per-class kept
// lists all live in shared memory, so the serial greedy loop of a class touches no
// global memory.  Slow path (rare overflow): keys sorted in the global workspace, boxes
// fetched on demand, kept boxes in the workspace.
template <typename Fetch>
__global__ void __launch_bounds__(kNmsThreads)
nms_image_kernel(NmsParams P, Fetch fetch, uint64_t* __restrict__ keys, uint64_t* __restrict__ merge,
                 float4* __restrict__ kept_ws, int* __restrict__ counts,
                 float4* __restrict__ out_boxes, float* __restrict__ out_a, float* __restrict__ out_b,
                 int32_t* __restrict__ out_valid) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t* s_sort = reinterpret_cast<uint64_t*>(s_raw);
    float4*   s_box  = reinterpret_cast<float4*>(s_raw + (size_t)P.smem_sort_slots * 8);
    uint16_t* s_kidx = reinterpret_cast<uint16_t*>(s_raw + (size_t)P.smem_sort_slots * 8 + (size_t)P.fast_slots * 16);
    __shared__ int s_seg_start[257];
    __shared__ int s_cnt[256], s_cur[256];
    __shared__ int s_mcount;
    __shared__ int s_all_matrix;       // every non-empty class takes the bit-matrix path (kept lists unused)
    __shared__ int s_mwords, s_dmax;   // 32-bit words of the bit matrix; largest class that uses it
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nwarps = kNmsThreads / 32;
    const int T = P.max_total;
    float4* ob = out_boxes + (size_t)b * T;
    float* oa = out_a + (size_t)b * T;
    float* oc = out_b + (size_t)b * T;

    int tslot = 0;
#define NMS_STAMP(tag) do { if (tid == 0) trace_stamp(P.trace, 0, tslot, tag); } while (0)
    NMS_STAMP(1);
    const int raw_count = counts[b];
    if (raw_count > P.cap) {                       // candidate list overflowed: report, emit zeros
        for (int r = tid; r < T; r += kNmsThreads) { ob[r] = make_float4(0, 0, 0, 0); oa[r] = 0.f; oc[r] = 0.f; }
        if (tid == 0) out_valid[b] = -1;
        return;
    }
    const int M = raw_count;

    // ---- 1. sort candidates: class asc, score desc, anchor asc -------------
    const int P1 = pow2_ceil(M);
    uint64_t* gk = keys + (size_t)b * P.key_stride;
    uint64_t* cand = (P1 <= P.smem_sort_slots) ? s_sort : gk;
    const bool fast = P1 <= P.fast_slots;
    const bool grouped = cand == s_sort && 2 * M <= P.smem_sort_slots;      // class-grouped counting sort (common case)
    for (int i = tid; i < 257; i += kNmsThreads) s_seg_start[i] = -1;
    for (int i = tid; i < 256; i += kNmsThreads) { s_cnt[i] = 0; }
    if (tid == 0) { s_mcount = 0; s_all_matrix = 0; }
    __syncthreads();
    if (grouped) {
        // a. class histogram -> segment starts; b. scatter into class segments (arbitrary order inside);
        // c. rank inside the own segment only (a class holds a few dozen candidates): sum n_c^2 compares instead
        //    of a CTA-wide sorting network
        uint64_t* tmp = s_sort + (P.smem_sort_slots >> 1);
        for (int i = tid; i < M; i += kNmsThreads) atomicAdd(&s_cnt[(int)(gk[i] >> 56)], 1);
        __syncthreads();
        NMS_STAMP(2);
        if (wid == 0) {                                // exclusive scan of the 256 class counts (8 per lane)
            int c8[8], t = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c8[q] = s_cnt[lane * 8 + q]; t += c8[q]; }
            int incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            int start = incl - t;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                s_cur[lane * 8 + q] = start;
                if (c8[q] > 0) s_seg_start[lane * 8 + q] = start;
                start += c8[q];
            }
        }
        __syncthreads();
        for (int i = tid; i < M; i += kNmsThreads) {
            const uint64_t key = gk[i];
            tmp[atomicAdd(&s_cur[(int)(key >> 56)], 1)] = key;
        }
        __syncthreads();
        NMS_STAMP(3);
        for (int i = tid; i < M; i += kNmsThreads) {
            const uint64_t key = tmp[i];
            const int c = (int)(key >> 56), start = s_seg_start[c], end = start + s_cnt[c];
            int rank = 0;
            for (int j = start; j < end; ++j) rank += (tmp[j] < key) ? 1 : 0;
            s_sort[start + rank] = key;
        }
        __syncthreads();
        NMS_STAMP(4);
        if (fast) {
            for (int i = tid; i < M; i += kNmsThreads) {
                const uint64_t key = cand[i];
                s_box[i] = nms_canonical(fetch(b, (int)(key & 0xFFFFFFu), (int)(key >> 56)));
            }
            // Suppression bit matrix, built by the WHOLE CTA (a warp per class leaves most of the CTA idle and runs
            // one dependent IoU chain per candidate): row li of class c holds, as 2*W 32-bit words (W = ceil(n_c/64)),
            // bit lj set iff lj > li and IoU(li, lj) > threshold.  Rows live in the half of the sort buffer the
            // counting sort no longer needs; s_cur[c] = first word of the class (-1: class keeps the serial path).
            if (wid == 0) {
                int w8[8], t = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = s_cnt[lane * 8 + q];
                    w8[q] = (n > 0 && n <= kMatrixMaxSeg) ? n * 2 * ((n + 63) >> 6) : 0;
                    t += w8[q];
                }
                int incl = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                int base = incl - t;
                bool all = true;
                int dmax = 0, used = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = s_cnt[lane * 8 + q];
                    const bool ok = w8[q] > 0 && base + w8[q] <= P.smem_sort_slots;     // u32 words in half the buffer
                    s_cur[lane * 8 + q] = ok ? base : -1;
                    all = all && (ok || n == 0);
                    if (ok) { dmax = max(dmax, n); used = base + w8[q]; }
                    base += w8[q];
                }
                all = __all_sync(0xffffffffu, all) && M <= (kNmsThreads / 128) * P.per_class;     // staging area: the u16 kept lists' bytes
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
                    used = max(used, __shfl_xor_sync(0xffffffffu, used, o));
                }
                if (lane == 0) { s_all_matrix = all ? 1 : 0; s_dmax = dmax; s_mwords = used; }
            }
            __syncthreads();
            NMS_STAMP(9);
            uint32_t* mask32 = reinterpret_cast<uint32_t*>(tmp);
            for (int t = tid; t < s_mwords; t += kNmsThreads) mask32[t] = 0u;
            __syncthreads();
            NMS_STAMP(10);
            // every (candidate, later candidate of its class) pair is an independent IoU test; the few overlapping pairs
            // set their bit atomically.  Classes of <= 64 candidates: one thread per candidate walks its later
            // partners (no index arithmetic); larger classes, or too few candidates to occupy the CTA
            // that way: the pairs are spread flat over all threads.
            const int D = s_dmax - 1;
            if (D < 64 && M * 2 > kNmsThreads) {
                for (int i = tid; i < M; i += kNmsThreads) {
                    const int c = (int)(cand[i] >> 56);
                    const int n = s_cnt[c], base = s_cur[c];
                    if (base < 0) continue;
                    const int start = s_seg_start[c], li = i - start;
                    const float4 bi = s_box[i];
                    uint32_t* row = mask32 + base + li * 2 * ((n + 63) >> 6);
                    for (int lj = li + 1; lj < n; ++lj)
                        if (nms_iou_gt_canonical(bi, s_box[start + lj], P.iou_thr)) atomicOr(&row[lj >> 5], 1u << (lj & 31));
                }
            } else {
                for (int t = tid; t < M * D; t += kNmsThreads) {
                    const int i = t / D, d = t - i * D + 1;
                    const int c = (int)(cand[i] >> 56);
                    const int n = s_cnt[c], base = s_cur[c];
                    const int start = s_seg_start[c], li = i - start, lj = li + d;
                    if (base < 0 || lj >= n) continue;
                    if (nms_iou_gt_canonical(s_box[i], s_box[start + lj], P.iou_thr))
                        atomicOr(&mask32[base + li * 2 * ((n + 63) >> 6) + (lj >> 5)], 1u << (lj & 31));
                }
            }
        }
    } else {
        if (cand == s_sort) {
            for (int i = tid; i < P1; i += kNmsThreads) s_sort[i] = (i < M) ? gk[i] : kPadKey;
        } else {
            for (int i = M + tid; i < P1; i += kNmsThreads) gk[i] = kPadKey;
        }
        __syncthreads();
        if (M > 1) bitonic_sort(cand, P1);

        // ---- 2. class segments; decode every candidate box once (fast path) ------
        for (int i = tid; i < M; i += kNmsThreads) {
            const uint64_t key = cand[i];
            const int c = (int)(key >> 56);
            if (i == 0 || (int)(cand[i - 1] >> 56) != c) s_seg_start[c] = i;
            if (fast) s_box[i] = nms_canonical(fetch(b, (int)(key & 0xFFFFFFu), c));
        }
    }
    __syncthreads();
    NMS_STAMP(5);

    // ---- 3. greedy suppression, one warp per class ---------------------------
    uint64_t* mk = merge + (size_t)b * P.merge_stride;
    for (int c = wid; c < P.L; c += nwarps) {
        const int start = s_seg_start[c];
        if (start < 0) continue;
        int nk = 0;
        const int seg = grouped ? s_cnt[c] : 0;          // segment length (known up front on the grouped path)
        if (fast && seg > 0 && s_cur[c] >= 0) {
            // Greedy scan over the class's rows of the bit matrix (built above by the whole CTA): candidate li survives
            // iff no EARLIER KEPT candidate has its bit set -- the same result as the candidate-by-candidate loop
            // below.  `removed` is replicated in every lane (warp-uniform control flow, broadcast row loads).
            const int W = (seg + 63) >> 6;                       // 1..4
            const uint64_t* rows = reinterpret_cast<const uint64_t*>(s_sort + (P.smem_sort_slots >> 1)) + (s_cur[c] >> 1);
            uint64_t removed[4] = {0ull, 0ull, 0ull, 0ull};
            int kept_n = 0;
            if (W == 1) {
                // <= 64 candidates: lane r holds rows r and r + 32; the scan pulls row i with a shuffle that does not
                // depend on `removed`, so only a shift / test / or chain is serial
                const uint64_t row0 = lane < seg ? rows[lane] : 0ull, row1 = lane + 32 < seg ? rows[lane + 32] : 0ull;
                uint64_t rem = 0ull;
                for (int i = 0; i < seg; ++i) {
                    const uint64_t row = __shfl_sync(0xffffffffu, i < 32 ? row0 : row1, i & 31);
                    const bool alive = !((rem >> i) & 1ull) && kept_n < P.per_class;
                    kept_n += alive ? 1 : 0;
                    rem |= alive ? row : (1ull << i);
                }
                removed[0] = rem;
            } else
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (w < W) {
                    const int nb = min(64, seg - (w << 6));
                    for (int bit = 0; bit < nb; ++bit) {
                        const int li = (w << 6) + bit;
                        const bool alive = !((removed[w] >> bit) & 1ull) && kept_n < P.per_class;
                        if (alive) {
                            ++kept_n;
#pragma unroll
                            for (int w2 = 0; w2 < 4; ++w2)
                                if (w2 >= w && w2 < W) removed[w2] |= rows[li * W + w2];
                        } else {
                            removed[w] |= 1ull << bit;           // suppressed, or beyond the per-class cap
                        }
                    }
                }
            }
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_mcount, kept_n);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint64_t* dst = s_all_matrix ? reinterpret_cast<uint64_t*>(s_kidx) : mk;
            int before = 0;                                      // kept candidates in earlier words
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (w < W) {
                    const int nb = min(64, seg - (w << 6));
                    const uint64_t valid = nb == 64 ? ~0ull : ((1ull << nb) - 1ull);
                    const uint64_t kept = ~removed[w] & valid;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int bit = lane + 32 * h;
                        if ((kept >> bit) & 1ull) {
                            const uint64_t key = cand[start + (w << 6) + bit];
                            const int slot = base + before + __popcll(kept & ((1ull << bit) - 1ull));
                            const uint32_t inv_score = (uint32_t)((key >> 24) & 0xFFFFFFFFu);
                            dst[slot] = ((uint64_t)inv_score << 32) | ((uint64_t)c << 24) | (uint32_t)(key & 0xFFFFFFu);
                        }
                    }
                    before += __popcll(kept);
                }
            }
        } else if (fast) {
            uint16_t* kidx = s_kidx + (size_t)wid * P.per_class;
            for (int i = start; i < M && nk < P.per_class; ++i) {
                const uint64_t key = cand[i];
                if ((int)(key >> 56) != c) break;
                const float4 box = s_box[i];
                bool sup = false;
                for (int j = lane; j < nk; j += 32) sup |= nms_iou_gt_canonical(box, s_box[kidx[j]], P.iou_thr);
                if (!__any_sync(0xffffffffu, sup)) {
                    if (lane == 0) {
                        kidx[nk] = (uint16_t)i;
                        const int slot = atomicAdd(&s_mcount, 1);
                        const uint32_t inv_score = (uint32_t)((key >> 24) & 0xFFFFFFFFu);
                        mk[slot] = ((uint64_t)inv_score << 32) | ((uint64_t)c << 24) | (uint32_t)(key & 0xFFFFFFu);
                    }
                    ++nk;
                    __syncwarp();
                }
            }
        } else {
            float4* kept = kept_ws + ((size_t)b * P.L + c) * P.per_class;
            for (int i = start; i < M && nk < P.per_class; ++i) {
                const uint64_t key = cand[i];
                if ((int)(key >> 56) != c) break;
                const int anchor = (int)(key & 0xFFFFFFu);
                const float4 box = fetch(b, anchor, c);
                bool sup = false;
                for (int j = lane; j < nk; j += 32) sup |= nms_iou_gt(box, __ldcg(kept + j), P.iou_thr);   // bypass L1
                if (!__any_sync(0xffffffffu, sup)) {
                    if (lane == 0) {
                        kept[nk] = box;
                        const int slot = atomicAdd(&s_mcount, 1);
                        const uint32_t inv_score = (uint32_t)((key >> 24) & 0xFFFFFFFFu);
                        mk[slot] = ((uint64_t)inv_score << 32) | ((uint64_t)c << 24) | (uint32_t)anchor;
                    }
                    ++nk;
                    __syncwarp();
                }
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    NMS_STAMP(6);

    // ---- 4. merge: score desc, class asc, anchor asc; first max_total --------
    const int K = s_mcount;
    const int P2 = pow2_ceil(K);
    uint64_t* ms = (P2 <= P.smem_sort_slots) ? s_sort : mk;
    if (ms == s_sort) {
        if (grouped && fast && s_all_matrix) {                  // kept keys were staged in shared memory
            const uint64_t* stage = reinterpret_cast<const uint64_t*>(s_kidx);
            for (int i = tid; i < P2; i += kNmsThreads) s_sort[i] = (i < K) ? stage[i] : kPadKey;
        } else {
            for (int i = tid; i < P2; i += kNmsThreads) s_sort[i] = (i < K) ? __ldcg(mk + i) : kPadKey;
        }
    } else {
        for (int i = K + tid; i < P2; i += kNmsThreads) mk[i] = kPadKey;
    }
    __syncthreads();
    if (K > 1) {
        if (ms == s_sort && K <= kRankSortMax && 2 * P2 <= P.smem_sort_slots) rank_sort(ms, s_sort + (P.smem_sort_slots >> 1), K);
        else bitonic_sort(ms, P2);
    }
    NMS_STAMP(7);
    const int V = min(K, T);
    for (int r = tid; r < T; r += kNmsThreads) {
        if (r < V) {
            uint64_t key = ms[r];
            int anchor = (int)(key & 0xFFFFFFu);
            int c = (int)((key >> 24) & 0xFFu);
            float score = unorder_bits(~(uint32_t)(key >> 32));
            float4 box = fetch(b, anchor, c);
            if (P.clip) {
                box.x = fminf(fmaxf(box.x, 0.f), 1.f); box.y = fminf(fmaxf(box.y, 0.f), 1.f);
                box.z = fminf(fmaxf(box.z, 0.f), 1.f); box.w = fminf(fmaxf(box.w, 0.f), 1.f);
            }
            ob[r] = box;
            if (P.labels_first) { oa[r] = (float)c; oc[r] = score; }
            else                { oa[r] = score;    oc[r] = (float)c; }
        } else {
            ob[r] = make_float4(0, 0, 0, 0); oa[r] = 0.f; oc[r] = 0.f;
        }
    }
    if (tid == 0) out_valid[b] = V;
    NMS_STAMP(8);
#undef NMS_STAMP
}

// ------------------------------------------------------------- host helpers --
struct NmsWs { int* counts; uint64_t* keys; uint64_t* merge; float4* kept; };

static int host_pow2_ceil(int64_t v) { int64_t p = 32; while (p < v) p <<= 1; return (int)p; }

static size_t nms_ws_layout(int B, int key_stride, int merge_stride, int L, int per_class, NmsWs* w, void* base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_c = take((size_t)B * 4);
    size_t o_k = take((size_t)B * key_stride * 8);
    size_t o_m = take((size_t)B * merge_stride * 8);
    size_t o_b = take((size_t)B * L * per_class * 16);
    if (w) {
        char* p = static_cast<char*>(base);
        w->counts = (int*)(p + o_c); w->keys = (uint64_t*)(p + o_k);
        w->merge = (uint64_t*)(p + o_m); w->kept = (float4*)(p + o_b);
    }
    return off;
}

struct NmsPlan { NmsParams p; size_t smem; size_t ws_bytes; };

static int make_plan(int B, int N, int L, int per_class, int max_total, int64_t cap64, NmsPlan* plan) {
    if (B < 0 || N < 1 || L < 1 || L > 255 || per_class < 1 || max_total < 1 || N >= (1 << 24) || B > 65535)
        return SSD_ERR_SHAPE;
    if (cap64 < 1 || cap64 > (int64_t)1 << 28) return SSD_ERR_SHAPE;
    if (per_class > N) per_class = N;               // a class can never keep more boxes than anchors
    if ((int64_t)L * per_class > (int64_t)1 << 26) return SSD_ERR_SHAPE;
    NmsParams& p = plan->p;
    p.N = N; p.L = L; p.per_class = per_class; p.max_total = max_total; p.cap = (int)cap64;
    p.trace = debug_trace_buffer();
    p.key_stride = host_pow2_ceil(cap64);
    p.merge_stride = host_pow2_ceil((int64_t)L * per_class);
    p.smem_sort_slots = min(4096, max(p.key_stride, p.merge_stride));
    // fast path: 8 B key + 16 B box per slot + u16 kept lists; 4096 slots with per_class = 200 is
    // 109 KB, i.e. two resident CTAs per SM
    p.fast_slots = (per_class <= 256) ? p.smem_sort_slots : 0;
    plan->smem = (size_t)p.smem_sort_slots * 8 + (size_t)p.fast_slots * 16 +
                 (p.fast_slots ? (size_t)(kNmsThreads / 32) * per_class * 2 : 0);
    plan->ws_bytes = nms_ws_layout(max(B, 1), p.key_stride, p.merge_stride, L, per_class, nullptr, nullptr);
    return SSD_OK;
}

template <typename Fetch>
static int run_image_pass(const NmsPlan& plan, Fetch fetch, const NmsWs& w, int B, float* d_boxes, float* d_a,
                          float* d_b, int32_t* d_valid, cudaStream_t st) {
    auto kern = nms_image_kernel<Fetch>;
    if (plan.smem > 40 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    kern<<<B, kNmsThreads, plan.smem, st>>>(plan.p, fetch, w.keys, w.merge, w.kept, w.counts,
                                            reinterpret_cast<float4*>(d_boxes), d_a, d_b, d_valid);
    SSD_CHECK_LAUNCH("nms_image_kernel");
    return SSD_OK;
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_softmax(const float* d_logits, int64_t rows, int L, float* d_probs, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_logits); SSD_REQUIRE_PTR(d_probs);
    SSD_REQUIRE(rows >= 0 && L >= 1, SSD_ERR_SHAPE, "ssd_softmax: bad shape rows=%lld L=%d", (long long)rows, L);
    if (rows == 0) return SSD_OK;
    size_t smem = ((size_t)kRowThreadsNms * L + 4) * sizeof(float);
    SSD_REQUIRE(smem <= 200 * 1024, SSD_ERR_UNSUPPORTED, "ssd_softmax: L=%d too large for row staging", L);
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (rows + kRowThreadsNms - 1) / kRowThreadsNms;
    int64_t gcap = (int64_t)sm_count() * 16;
    int grid = (int)(blocks < gcap ? blocks : gcap);
    softmax_kernel<<<grid, kRowThreadsNms, smem, as_stream(stream)>>>(d_logits, rows, L, d_probs);
    SSD_CHECK_LAUNCH("softmax_kernel");
    return SSD_OK;
}

extern "C" size_t ssd_decode_nms_workspace_bytes(int B, int N, int L, int max_total_size, int max_candidates) {
    NmsPlan plan;
    int64_t cap = max_candidates > 0 ? max_candidates : N;
    if (make_plan(B, N, L, max_total_size, max_total_size, cap, &plan) != SSD_OK) return 0;
    return plan.ws_bytes;
}

extern "C" int ssd_decode_nms(const float* d_priors, const float* d_pred_deltas, const float* d_pred_labels,
                              int B, int N, int L, const float* h_variances, int from_logits,
                              float score_threshold, float iou_threshold, int max_total_size, int max_candidates,
                              float* d_boxes, float* d_labels, float* d_scores, int32_t* d_valid,
                              void* d_workspace, size_t workspace_bytes, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_priors); SSD_REQUIRE_PTR(d_pred_deltas); SSD_REQUIRE_PTR(d_pred_labels);
    SSD_REQUIRE_PTR(h_variances); SSD_REQUIRE_PTR(d_boxes); SSD_REQUIRE_PTR(d_labels);
    SSD_REQUIRE_PTR(d_scores); SSD_REQUIRE_PTR(d_valid);
    NmsPlan plan;
    int64_t cap = max_candidates > 0 ? max_candidates : N;
    SSD_REQUIRE(make_plan(B, N, L, max_total_size, max_total_size, cap, &plan) == SSD_OK, SSD_ERR_SHAPE,
                "ssd_decode_nms: bad shape B=%d N=%d L=%d max_total=%d cap=%lld", B, N, L, max_total_size,
                (long long)cap);
    if (B == 0) return SSD_OK;
    SSD_REQUIRE_PTR(d_workspace);
    SSD_REQUIRE(workspace_bytes >= plan.ws_bytes, SSD_ERR_WORKSPACE,
                "ssd_decode_nms: workspace %zu < required %zu bytes", workspace_bytes, plan.ws_bytes);
    plan.p.iou_thr = iou_threshold; plan.p.clip = 1; plan.p.labels_first = 1;
    NmsWs w;
    nms_ws_layout(B, plan.p.key_stride, plan.p.merge_stride, L, plan.p.per_class, &w, d_workspace);
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(w.counts, 0, (size_t)B * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "ssd_decode_nms: memset");

    size_t smem = ((size_t)kRowThreadsNms * L + 4) * sizeof(float);
    SSD_REQUIRE(smem <= 200 * 1024, SSD_ERR_UNSUPPORTED, "ssd_decode_nms: L=%d too large", L);
    dim3 grid(ceil_div(N, kRowThreadsNms), B);
    auto launch = [&](auto kern) {
        if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kRowThreadsNms, smem, st>>>(d_pred_labels, N, L, score_threshold, plan.p.cap, w.keys,
                                                 plan.p.key_stride, w.counts);
    };
    if (from_logits) { if (L == 21) launch(nms_candidates_kernel<true, true, 21>); else launch(nms_candidates_kernel<true, true, 0>); }
    else launch(nms_candidates_kernel<true, false, 0>);
    SSD_CHECK_LAUNCH("nms_candidates_kernel");

    DecodeFetch fetch{reinterpret_cast<const float4*>(d_priors), reinterpret_cast<const float4*>(d_pred_deltas),
                      make_float4(h_variances[0], h_variances[1], h_variances[2], h_variances[3]), N};
    return run_image_pass(plan, fetch, w, B, d_boxes, d_labels, d_scores, d_valid, st);
}

extern "C" size_t ssd_combined_nms_workspace_bytes(int B, int N, int L, int max_output_size_per_class,
                                                   int max_total_size, int max_candidates) {
    NmsPlan plan;
    int64_t cap = max_candidates > 0 ? max_candidates : (int64_t)N * L;
    if (make_plan(B, N, L, max_output_size_per_class, max_total_size, cap, &plan) != SSD_OK) return 0;
    return plan.ws_bytes;
}

extern "C" int ssd_combined_nms(const float* d_boxes, const float* d_scores, int B, int N, int q, int L,
                                int max_output_size_per_class, int max_total_size,
                                float iou_threshold, float score_threshold, int clip_boxes, int max_candidates,
                                float* d_out_boxes, float* d_out_scores, float* d_out_classes, int32_t* d_valid,
                                void* d_workspace, size_t workspace_bytes, ssd_stream_t stream) {
    SSD_REQUIRE_PTR(d_boxes); SSD_REQUIRE_PTR(d_scores); SSD_REQUIRE_PTR(d_out_boxes);
    SSD_REQUIRE_PTR(d_out_scores); SSD_REQUIRE_PTR(d_out_classes); SSD_REQUIRE_PTR(d_valid);
    SSD_REQUIRE(q == 1 || q == L, SSD_ERR_SHAPE, "ssd_combined_nms: q=%d must be 1 or L=%d", q, L);
    NmsPlan plan;
    int64_t cap = max_candidates > 0 ? max_candidates : (int64_t)N * L;
    SSD_REQUIRE(make_plan(B, N, L, max_output_size_per_class, max_total_size, cap, &plan) == SSD_OK, SSD_ERR_SHAPE,
                "ssd_combined_nms: bad shape B=%d N=%d L=%d per_class=%d max_total=%d cap=%lld", B, N, L,
                max_output_size_per_class, max_total_size, (long long)cap);
    if (B == 0) return SSD_OK;
    SSD_REQUIRE_PTR(d_workspace);
    SSD_REQUIRE(workspace_bytes >= plan.ws_bytes, SSD_ERR_WORKSPACE,
                "ssd_combined_nms: workspace %zu < required %zu bytes", workspace_bytes, plan.ws_bytes);
    plan.p.iou_thr = iou_threshold; plan.p.clip = clip_boxes ? 1 : 0; plan.p.labels_first = 0;
    NmsWs w;
    nms_ws_layout(B, plan.p.key_stride, plan.p.merge_stride, L, plan.p.per_class, &w, d_workspace);
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(w.counts, 0, (size_t)B * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "ssd_combined_nms: memset");

    size_t smem = ((size_t)kRowThreadsNms * L + 4) * sizeof(float);
    SSD_REQUIRE(smem <= 200 * 1024, SSD_ERR_UNSUPPORTED, "ssd_combined_nms: L=%d too large", L);
    dim3 grid(ceil_div(N, kRowThreadsNms), B);
    auto kern = nms_candidates_kernel<false, false, 0>;
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, kRowThreadsNms, smem, st>>>(d_scores, N, L, score_threshold, plan.p.cap, w.keys,
                                             plan.p.key_stride, w.counts);
    SSD_CHECK_LAUNCH("nms_candidates_kernel");
    DirectFetch fetch{reinterpret_cast<const float4*>(d_boxes), N, q};
    return run_image_pass(plan, fetch, w, B, d_out_boxes, d_out_scores, d_out_classes, d_valid, st);
}
