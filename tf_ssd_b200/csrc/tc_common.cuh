// Shared pieces of the tcgen05 / TMEM / TMA kernels of libssd_b200 (sm_100a): tile constants, the parameter block,
// PTX wrappers (mbarrier, TMA loads / stores, UMMA descriptors, tcgen05.mma / ld / commit) and the epilogue helpers.
#pragma once

#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace ssd {

constexpr int TC_BM = 128;           // UMMA M (one CTA, cta_group::1)
constexpr int TC_BK = 64;            // 64 fp16 = 128 B = one swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_PAIR_STAGES = 6;     // operand ring of the CTA-pair kernel (half the weight bytes per stage)
constexpr int TC_EPI_WARPS = 8;      // two warps per TMEM lane quarter (they split the column chunks)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..9: epilogue
// fused depthwise -> 1x1 kernel: roles on warpgroup boundaries (setmaxnreg works per warpgroup) -- warpgroup 0: warp 0 TMA,
// warp 1 MMA, warps 2-3 idle; warpgroups 1-2 (warps 4..11): epilogue; warpgroups 3-4 (warps 12..19): depthwise
constexpr int TC_DW_EPI_WARP0 = 4;
constexpr int TC_DW_WARP0 = TC_DW_EPI_WARP0 + TC_EPI_WARPS;
constexpr int TC_DW_WARPS = 8;
constexpr int TC_DW_ROWS = (TC_BM + 4 * TC_DW_WARPS - 1) / (4 * TC_DW_WARPS);   // tile rows per depthwise thread (4)
constexpr int TC_THREADS_DW = 32 * (TC_DW_WARP0 + TC_DW_WARPS);                  // 640
// setmaxnreg moves registers inside the CTA's own pool (640 threads x 96 launch registers): budgets sum to <= 5 x 96
constexpr int TC_DW_REGS_LAUNCH = 96, TC_DW_REGS_SCHED = 56, TC_DW_REGS_EPI = 80, TC_DW_REGS_DW = 128;
static_assert(TC_DW_REGS_SCHED + 2 * TC_DW_REGS_EPI + 2 * TC_DW_REGS_DW <= 5 * TC_DW_REGS_LAUNCH, "CTA register pool");
static_assert(TC_THREADS_DW * TC_DW_REGS_LAUNCH <= 65536 && TC_THREADS_DW * (TC_DW_REGS_LAUNCH + 8) > 65536, "launch registers = 96");
constexpr int TC_DW_ASTAGES = 2;     // (A, B) operand ring of the fused kernel
constexpr int TC_CHUNK = 32;         // epilogue column chunk per warp (fp16: 64 B per row)
constexpr int TC_OUT_TILE = TC_BM * 128;             // one output staging tile: 128 rows x 64 fp16 channels, 128B-swizzled
constexpr int TC_OUT_BYTES = 2 * TC_OUT_TILE + 256 * 4;   // two tiles (double buffer) + the N tile's bias values

struct TcParams {
    int mode4d;                      // 0: A is a 2-D [M, K] matrix (1x1 conv); 1: 4-D im2col boxes
    int M;                           // B*Ho*Wo
    int B, Ho, Wo, HoWo;
    int bw, bh, bb;                  // output-pixel box of one M tile (4-D mode), bw*bh*bb <= 128
    int tiles_w, tiles_h;            // tiles per image row / column
    int Cin, KW, dil, pad_t, pad_l, stride;
    int kb_per_tap;                  // ceil(Cin / 64)
    int n_kblocks;                   // taps * kb_per_tap
    int kb_per_split;                // k-blocks handled by one blockIdx.z
    int Cout, BN;
    uint32_t a_bytes, b_bytes;       // TMA transaction bytes per stage
    uint32_t idesc;                  // tcgen05 instruction descriptor
    uint32_t tmem_cols;              // allocated columns = 2 accumulators
    uint32_t acc_cols;               // column offset between the two accumulators
    int tiles_m, tiles_n, n_tiles;   // n_tiles = tiles_m * tiles_n * splits
    // epilogue
    const float* bias; const __half* res; void* out0; void* out1;
    int act, out_f32, split;
    long long img0, pix0, img1, pix1;
    float* partial;                  // split-K workspace [splits][tiles_m*128][ldp] or nullptr
    int splits, ldp;
    int tma_store;                   // fp16 single-segment output written with TMA stores (map_o is valid)
    int stages;                      // operand ring depth (2..TC_STAGES): shallow-K layers trade depth for 2 CTAs per SM
    // fused depthwise 3x3 (+ bias + activation) producing the A operand of a 1x1 convolution (DW kernel variant)
    const uint4* dw_x; const uint4* dw_w; const float* dw_bias;
    int dw_H, dw_W, dw_C8, dw_stride, dw_pad_t, dw_pad_l, dw_act;
    int dw_pw, dw_ph;                // input patch of one tile: pw x ph x bb positions of 64 channels
    uint32_t dw_patch_bytes;         // TMA transaction bytes of one patch
    uint32_t dw_patch_stage;         // patch stage size in shared memory (1024-byte multiple)
    int dw_pstages;                  // patch ring depth
    int dw_quad;                     // stride 1 and bw % 4 == 0: a thread owns 4 horizontally adjacent pixels (sliding window)
    int dw_pair;                     // <= 32 channels, stride 1, even bw: 4 live chunks x 2 adjacent pixels per thread
    unsigned long long* trace;       // debug (ssd_debug_trace): per-role stamps of CTA 0, else nullptr
};

// ------------------------------------------------------------------ PTX glue --
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
// The loaded registers are operands of the wait so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
// TMA stores (shared -> global, bulk async-group completion); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory"); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == SSD_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == SSD_ACT_RELU6) return fminf(fmaxf(v, 0.0f), 6.0f);
    return v;
}

// Output location of tile row r: returns false when the row is padding of the tile.
__device__ __forceinline__ bool tc_row_to_pixel(const TcParams& p, int tile, int r, int& b, int& pix) {
    if (!p.mode4d) {
        int m = tile * TC_BM + r;
        if (m >= p.M) return false;
        b = m / p.HoWo;
        pix = m - b * p.HoWo;
        return true;
    }
    const int per_img = p.tiles_w * p.tiles_h;
    const int tb = tile / per_img, tr = tile - tb * per_img;
    const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
    const int dx = r % p.bw, q = r / p.bw, dy = q % p.bh, db = q / p.bh;
    if (db >= p.bb) return false;
    b = tb * p.bb + db;
    const int oy = th * p.bh + dy, ox = tw * p.bw + dx;
    if (b >= p.B || oy >= p.Ho || ox >= p.Wo) return false;
    pix = oy * p.Wo + ox;
    return true;
}

// Final epilogue for one element (shared by the fused epilogue's generic path and the split-K reduction).
__device__ __forceinline__ void tc_store_one(const TcParams& p, int b, int pix, int n, float v) {
    if (p.bias) v += __ldg(p.bias + n);
    v = apply_act(v, p.act);
    const bool seg1 = n >= p.split;
    const size_t off = seg1 ? (size_t)b * p.img1 + (size_t)pix * p.pix1 + (n - p.split)
                            : (size_t)b * p.img0 + (size_t)pix * p.pix0 + n;
    void* base = seg1 ? p.out1 : p.out0;
    if (p.res && !seg1) v += __half2float(p.res[off]);
    if (p.out_f32) reinterpret_cast<float*>(base)[off] = v;
    else reinterpret_cast<__half*>(base)[off] = __float2half_rn(v);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}


// Encoded tensor maps, cached per (pointer, geometry) -- defined in conv_tcgen05.cu.
// swizzle_bytes: 128 (the tcgen05 operand tiles), 64 or 32 (the 32-byte-row slices of conv_irblock_mma.cu).
int cached_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, const uint32_t* elem_strides = nullptr, int swizzle_bytes = 128);

}  // namespace ssd
