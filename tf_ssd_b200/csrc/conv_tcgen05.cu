// Blackwell-native convolution for sm_100a: implicit GEMM on the 5th-generation
// tensor cores.
//
//   * operands staged by TMA (cp.async.bulk.tensor) into 128-byte-swizzled
//     shared-memory tiles; the im2col gather of a 3x3 / dilated convolution is
//     done BY the TMA unit: a 4-D tensor map over the NHWC activation is read
//     with the tap's (dy, dx) added to the box coordinates, and out-of-bounds
//     (padding) elements arrive as zeros;
//   * tcgen05.mma (kind::f16, M = 128, N = 16..256) issued by one elected
//     thread, fp32 accumulators in tensor memory (TMEM);
//   * PERSISTENT CTAs (one per SM) loop over output tiles; a 4-stage mbarrier
//     producer/consumer ring (TMA warp -> MMA warp) runs ahead across tiles and
//     the accumulator is double-buffered in TMEM (tmem_full / tmem_empty
//     barriers), so the epilogue of tile i overlaps the loads and MMAs of
//     tile i+1;
//   * epilogue (8 warps): tcgen05.ld -> registers -> bias / BatchNorm-folded
//     bias, ReLU / ReLU6, residual add -> fp16 -> per-warp shared-memory staging
//     -> fully coalesced 16-byte global stores; fp32 / two-segment strided
//     outputs (the multibox head writes straight into the concatenated tensors)
//     take a generic path;
//   * split-K across blockIdx.z for the head convolutions (tiny M, K up to
//     11 520): fp32 partials in a workspace + a deterministic fixed-order
//     reduction kernel that applies the epilogue.
//
// Handles every Conv2D of the two graphs (1x1, 3x3, dilation 6, SAME / VALID, stride 1 and 2 -- a strided
// convolution uses a tensor map whose W/H element strides equal the convolution stride, so the im2col box still
// arrives as bw x bh output pixels).  conv_igemm.cu (mma.sync) remains as the fallback for unsupported shapes.

#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace ssd {

// ------------------------------------------------------------------- kernel --
__device__ __forceinline__ void conv_tcgen05_body(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_o,
                                                  const TcParams& p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_stage = TC_BM * TC_BK * 2;                 // 16 KB
    const uint32_t b_stage = (uint32_t)p.BN * TC_BK * 2;
    unsigned char* sA = smem;
    unsigned char* sB = smem + p.stages * a_stage;
    unsigned char* sOut = sB + p.stages * b_stage;               // [2][128 rows][128 B] swizzled output tiles (1024-aligned)
    float* sBias = reinterpret_cast<float*>(sOut + 2 * TC_OUT_TILE);     // [256]: bias of the current N tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + TC_OUT_BYTES);
    // bars: full[S] | empty[S] | tmem_full[2] | tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_addr(bars), bar_empty = smem_addr(bars + TC_STAGES);
    const uint32_t bar_tfull = smem_addr(bars + 2 * TC_STAGES), bar_tempty = smem_addr(bars + 2 * TC_STAGES + 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (p.tma_store) tma_prefetch_desc(&map_o);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);                      // one arrive.expect_tx (+ TMA bytes)
            mbar_init(bar_empty + 8 * s, 1);                     // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);                     // one tcgen05.commit
            mbar_init(bar_tempty + 8 * a, TC_EPI_WARPS);         // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                             // this warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_addr(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_mn = p.tiles_m * p.tiles_n;
    pdl_wait();                                                  // predecessors' outputs are visible from here on

    if (warp == 0) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            int it = 0, s = 0;                                   // k-block counter across all tiles of this CTA, ring slot
            uint32_t ph = 0;                                     // parity of the ring's wrap count
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const int z = t / tiles_mn, tm_ = (t - z * tiles_mn) / p.tiles_n, tn = t - z * tiles_mn - tm_ * p.tiles_n;
                const int kb0 = z * p.kb_per_split, kb1 = min(p.n_kblocks, kb0 + p.kb_per_split);
                int b0 = 0, oy0 = 0, ox0 = 0;
                if (p.mode4d) {
                    const int per_img = p.tiles_w * p.tiles_h;
                    const int tb = tm_ / per_img, tr = tm_ - tb * per_img;
                    const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
                    b0 = tb * p.bb; oy0 = th * p.bh; ox0 = tw * p.bw;
                }
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    if (it >= p.stages) mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * TC_BK;
                    mbar_expect_tx(bar_full + 8 * s, p.a_bytes + p.b_bytes);
                    const uint32_t dst_a = smem_addr(sA + (size_t)s * a_stage), dst_b = smem_addr(sB + (size_t)s * b_stage);
                    if (p.mode4d) {
                        const int ky = tap / p.KW, kx = tap - ky * p.KW;
                        // strided convolutions: the tensor map traverses W and H with element stride = conv stride,
                        // so the box still lands as bw x bh output pixels
                        tma_load_4d(dst_a, &map_a, bar_full + 8 * s, c0, ox0 * p.stride + kx * p.dil - p.pad_l,
                                    oy0 * p.stride + ky * p.dil - p.pad_t, b0);
                    } else {
                        tma_load_2d(dst_a, &map_a, bar_full + 8 * s, c0, tm_ * TC_BM);
                    }
                    tma_load_2d(dst_b, &map_b, bar_full + 8 * s, tap * p.Cin + c0, tn * p.BN);
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        if (lane == 0) {
            int j = 0, s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++j) {
                const int z = t / tiles_mn;
                const int kb0 = z * p.kb_per_split, kb1 = min(p.n_kblocks, kb0 + p.kb_per_split);
                const int a = j & 1;
                if (j >= 2) mbar_wait(bar_tempty + 8 * a, ((j >> 1) - 1) & 1);      // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)a * p.acc_cols;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_full + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc_sw128(smem_addr(sA + (size_t)s * a_stage));
                    const uint64_t db = umma_desc_sw128(smem_addr(sB + (size_t)s * b_stage));
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)         // UMMA_K = 16: +32 bytes along K inside the swizzle row
                        umma_f16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (kb > kb0) || k > 0);
                    umma_commit(bar_empty + 8 * s);              // frees the stage when these MMAs retire
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(bar_tfull + 8 * a);                  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
        const int ew = warp - 2;
        const int q = warp & 3;                                  // tcgen05.ld: warp w may touch lanes 32*(w%4)..+31
        const int half = ew >> 2;                                // which 32-column half of a 64-column group this warp takes
        const int r = q * 32 + lane;
        const bool elected = ew == 0 && lane == 0;               // issues / retires the TMA stores of this CTA
        int bias_tn = -1;                                        // N tile whose bias sits in sBias
        const float act_lo = p.act == SSD_ACT_NONE ? -__int_as_float(0x7f800000) : 0.0f;
        const float act_hi = p.act == SSD_ACT_RELU6 ? 6.0f : __int_as_float(0x7f800000);
        int j = 0;
        uint32_t n_groups = 0;                                   // output groups staged so far (selects the buffer)
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++j) {
            const int z = t / tiles_mn, tm_ = (t - z * tiles_mn) / p.tiles_n, tn = t - z * tiles_mn - tm_ * p.tiles_n;
            const int n0 = tn * p.BN;
            const int a = j & 1;
            mbar_wait(bar_tfull + 8 * a, (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int b = 0, pix = 0;
            const bool row_ok = tc_row_to_pixel(p, tm_, r, b, pix);
            const long long row_off = (long long)b * p.img0 + (long long)pix * p.pix0;     // element offset of the row (segment 0)
            const uint32_t trow = tmem_base + (uint32_t)a * p.acc_cols + ((uint32_t)(q * 32) << 16);
            if (p.tma_store) {
                // ---- fp16 output through shared memory + TMA: 64-channel groups, double-buffered staging tile.
                // Thread r owns tile row r; its 32 channels (64 B) go to 16-byte chunks (half*4 + h) ^ (r & 7) of
                // the row's 128 bytes -- the SWIZZLE_128B layout of map_o, conflict-free for the warp.
                int ox0 = 0, oy0 = 0, b0 = 0;
                if (p.mode4d) {
                    const int per_img = p.tiles_w * p.tiles_h;
                    const int tb = tm_ / per_img, tr = tm_ - tb * per_img;
                    const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
                    b0 = tb * p.bb; oy0 = th * p.bh; ox0 = tw * p.bw;
                }
                if (bias_tn != tn) {                             // (re)load this N tile's bias: visible after the next barrier;
                    const int i = ew * 32 + lane;                // the previous tile's readers are all past their last barrier
                    sBias[i] = (p.bias && i < p.BN && n0 + i < p.Cout) ? __ldg(p.bias + n0 + i) : 0.0f;
                    bias_tn = tn;
                }
                for (int g0 = 0; g0 < p.BN && n0 + g0 < p.Cout; g0 += 64, ++n_groups) {
                    unsigned char* buf = sOut + (n_groups & 1u) * TC_OUT_TILE;
                    const int c0 = g0 + half * TC_CHUNK;
                    const bool mine = c0 < p.BN && n0 + c0 < p.Cout;   // warp-uniform
                    uint32_t acc[TC_CHUNK];
                    if (mine) tmem_ld32(trow + (uint32_t)c0, acc);     // in flight across the buffer hand-over below
                    if (elected) bulk_wait_read<1>();            // the store that last read this buffer has drained it
                    epi_barrier();
                    if (mine) {
                        const int n = n0 + c0;
                        const int ncols = min(TC_CHUNK, p.Cout - n);
                        const float* sbias = sBias + c0;
                        tmem_ld_wait(acc);
#pragma unroll
                        for (int h = 0; h < TC_CHUNK / 8; ++h) {
                            const float4 b0v = *reinterpret_cast<const float4*>(sbias + h * 8);
                            const float4 b1v = *reinterpret_cast<const float4*>(sbias + h * 8 + 4);
                            const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                v[e] = fminf(fmaxf(__uint_as_float(acc[h * 8 + e]) + bb[e], act_lo), act_hi);
                            if (p.res && row_ok && h * 8 < ncols) {
                                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + row_off + n + h * 8));
                                const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __half22float2(rh[e]);
                                    v[2 * e] += f.x; v[2 * e + 1] += f.y;
                                }
                            }
                            uint4 o;
                            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                            for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                            *reinterpret_cast<uint4*>(buf + r * 128 + (((half * 4 + h) ^ (r & 7)) << 4)) = o;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    epi_barrier();
                    if (elected) {
                        if (p.mode4d) tma_store_4d(&map_o, smem_addr(buf), n0 + g0, ox0, oy0, b0);
                        else          tma_store_2d(&map_o, smem_addr(buf), n0 + g0, tm_ * TC_BM);
                        bulk_commit();
                    }
                }
            } else {
                for (int c0 = half * TC_CHUNK; c0 < p.BN; c0 += 2 * TC_CHUNK) {
                    if (n0 + c0 >= p.Cout) break;                // warp-uniform
                    uint32_t acc[TC_CHUNK];
                    tmem_ld32(trow + (uint32_t)c0, acc);
                    tmem_ld_wait(acc);
                    const int ncols = min(min(TC_CHUNK, p.BN - c0), p.Cout - (n0 + c0));    // valid columns of this chunk
                    if (p.partial) {
                        if (row_ok) {
                            float* dst = p.partial + ((size_t)z * p.tiles_m * TC_BM + (size_t)tm_ * TC_BM + r) * p.ldp + n0 + c0;
#pragma unroll
                            for (int jj = 0; jj < TC_CHUNK; jj += 4)
                                if (jj < p.BN - c0)
                                    *reinterpret_cast<float4*>(dst + jj) = make_float4(__uint_as_float(acc[jj]), __uint_as_float(acc[jj + 1]),
                                                                                       __uint_as_float(acc[jj + 2]), __uint_as_float(acc[jj + 3]));
                        }
                    } else if (row_ok) {
#pragma unroll
                        for (int jj = 0; jj < TC_CHUNK; ++jj)
                            if (jj < ncols) tc_store_one(p, b, pix, n0 + c0 + jj, __uint_as_float(acc[jj]));
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * a);      // this warp is done reading accumulator a
        }
        if (elected) bulk_wait_all();                            // staged tiles must outlive their stores
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

__global__ void __launch_bounds__(TC_THREADS, 2)     // <= 102 registers: two CTAs may share an SM (shallow-K layers)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_o, const __grid_constant__ TcParams p) {
    conv_tcgen05_body(map_a, map_b, map_o, p);
}
// ======================================================================================
// The same implicit GEMM on a CTA PAIR (cta_group::2): the VGG16 body.  Two CTAs of a cluster (the two SMs of a TPC)
// execute ONE tcgen05.mma of M = 256: each CTA stages its own 128 output pixels of A and only HALF of the weight tile
// (BN/2 rows), the leader's MMA reads both halves -- per CTA and k-block 16 KB + BN/2 x 128 B instead of 16 KB +
// BN x 128 B, i.e. two thirds of the operand traffic that bounds the 3x3 layers (im2col re-reads the activation nine
// times and every M tile re-reads the weights), and a deeper ring in the same shared memory.
//   * both CTAs issue their TMA loads with .cta_group::2 and signal the LEADER's "full" barrier (expect_tx there counts
//     the bytes of both); the leader's single thread issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts the
//     "stage empty" / "accumulator full" arrivals to the barriers of both CTAs;
//   * every CTA drains its own 128 TMEM lanes; the peer's epilogue warps arrive remotely (mapa) on the leader's
//     "accumulator empty" barrier;
//   * cluster barriers after the barrier initialisation and before the TMEM deallocation.
// fp16 single-segment outputs (TMA-store epilogue) without split-K only; an odd number of M tiles leaves a phantom
// tile whose loads are zero-filled and whose stores are clipped by the tensor maps.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {          // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_o, const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_stage = TC_BM * TC_BK * 2;                 // 16 KB: this CTA's 128 pixels
    const uint32_t b_stage = (uint32_t)(p.BN / 2) * TC_BK * 2;  // this CTA's half of the weight tile
    unsigned char* sA = smem;
    unsigned char* sB = smem + p.stages * a_stage;
    unsigned char* sOut = sB + p.stages * b_stage;
    float* sBias = reinterpret_cast<float*>(sOut + 2 * TC_OUT_TILE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + TC_OUT_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_PAIR_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t bar_full = smem_addr(bars), bar_empty = smem_addr(bars + TC_PAIR_STAGES);
    const uint32_t bar_tfull = smem_addr(bars + 2 * TC_PAIR_STAGES), bar_tempty = smem_addr(bars + 2 * TC_PAIR_STAGES + 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_o);
        for (int s = 0; s < TC_PAIR_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);                      // leader: one arrive.expect_tx (+ the TMA bytes of both CTAs)
            mbar_init(bar_empty + 8 * s, 1);                     // one multicast tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);                     // one multicast tcgen05.commit
            mbar_init(bar_tempty + 8 * a, 2 * TC_EPI_WARPS);     // leader: the epilogue warps of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                             // one warp of EACH CTA takes part in the pair allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_addr(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_barrier();                                           // barriers of both CTAs are initialised, TMEM is allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int pair_m = (p.tiles_m + 1) >> 1;                     // pairs of M tiles
    const int n_pair_tiles = pair_m * p.tiles_n;
    const int n_pairs = (int)gridDim.x >> 1, pair_id = (int)blockIdx.x >> 1;
    const uint32_t full_leader = mapa_rank(bar_full, 0);         // the leader's "full" barriers in cluster address space

    if (warp == 0) {
        // ===================== TMA producer (one elected lane of each CTA) =====================
        if (lane == 0) {
            int it = 0, s = 0;
            uint32_t ph = 0;
            for (int t = pair_id; t < n_pair_tiles; t += n_pairs) {
                const int pm = t / p.tiles_n, tn = t - pm * p.tiles_n;
                const int tm_ = 2 * pm + (int)rank;
                int b0 = 0, oy0 = 0, ox0 = 0;
                if (p.mode4d) {
                    const int per_img = p.tiles_w * p.tiles_h;
                    const int tb = tm_ / per_img, tr = tm_ - tb * per_img;
                    const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
                    b0 = tb * p.bb; oy0 = th * p.bh; ox0 = tw * p.bw;
                }
                for (int kb = 0; kb < p.n_kblocks; ++kb, ++it) {
                    if (it >= p.stages) mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * TC_BK;
                    if (leader) mbar_expect_tx(bar_full + 8 * s, 2u * (p.a_bytes + p.b_bytes));    // p.b_bytes: one half
                    const uint32_t dst_a = smem_addr(sA + (size_t)s * a_stage), dst_b = smem_addr(sB + (size_t)s * b_stage);
                    const uint32_t bar = full_leader + 8 * s;
                    if (p.mode4d) {
                        const int ky = tap / p.KW, kx = tap - ky * p.KW;
                        tma2_load_4d(dst_a, &map_a, bar, c0, ox0 * p.stride + kx * p.dil - p.pad_l,
                                     oy0 * p.stride + ky * p.dil - p.pad_t, b0);
                    } else {
                        tma2_load_2d(dst_a, &map_a, bar, c0, tm_ * TC_BM);
                    }
                    tma2_load_2d(dst_b, &map_b, bar, tap * p.Cin + c0, tn * p.BN + (int)rank * (p.BN / 2));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: the leader's elected lane, M = 256 over both CTAs =====================
        if (leader && lane == 0) {
            int j = 0, s = 0;
            uint32_t ph = 0;
            for (int t = pair_id; t < n_pair_tiles; t += n_pairs, ++j) {
                const int a = j & 1;
                if (j >= 2) mbar_wait(bar_tempty + 8 * a, ((j >> 1) - 1) & 1);      // both epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)a * p.acc_cols;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc_sw128(smem_addr(sA + (size_t)s * a_stage));
                    const uint64_t db = umma_desc_sw128(smem_addr(sB + (size_t)s * b_stage));
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        umma2_f16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (kb > 0) || k > 0);
                    umma2_commit(bar_empty + 8 * s);             // frees the stage in both CTAs
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                umma2_commit(bar_tfull + 8 * a);                 // accumulator complete, in both CTAs
            }
        }
    } else {
        // ===================== epilogue: as conv_tcgen05_kernel's TMA-store path, on this CTA's 128 rows =====================
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int r = q * 32 + lane;
        const bool elected = ew == 0 && lane == 0;
        int bias_tn = -1;
        const float act_lo = p.act == SSD_ACT_NONE ? -__int_as_float(0x7f800000) : 0.0f;
        const float act_hi = p.act == SSD_ACT_RELU6 ? 6.0f : __int_as_float(0x7f800000);
        const uint32_t tempty_leader = mapa_rank(bar_tempty, 0);
        int j = 0;
        uint32_t n_groups = 0;
        for (int t = pair_id; t < n_pair_tiles; t += n_pairs, ++j) {
            const int pm = t / p.tiles_n, tn = t - pm * p.tiles_n;
            const int tm_ = 2 * pm + (int)rank;
            const int n0 = tn * p.BN;
            const int a = j & 1;
            mbar_wait(bar_tfull + 8 * a, (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int b = 0, pix = 0;
            const bool row_ok = tc_row_to_pixel(p, tm_, r, b, pix);
            const long long row_off = (long long)b * p.img0 + (long long)pix * p.pix0;
            const uint32_t trow = tmem_base + (uint32_t)a * p.acc_cols + ((uint32_t)(q * 32) << 16);
            int ox0 = 0, oy0 = 0, b0 = 0;
            if (p.mode4d) {
                const int per_img = p.tiles_w * p.tiles_h;
                const int tb = tm_ / per_img, tr = tm_ - tb * per_img;
                const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
                b0 = tb * p.bb; oy0 = th * p.bh; ox0 = tw * p.bw;
            }
            if (bias_tn != tn) {
                const int i = ew * 32 + lane;
                sBias[i] = (p.bias && i < p.BN && n0 + i < p.Cout) ? __ldg(p.bias + n0 + i) : 0.0f;
                bias_tn = tn;
            }
            for (int g0 = 0; g0 < p.BN && n0 + g0 < p.Cout; g0 += 64, ++n_groups) {
                unsigned char* buf = sOut + (n_groups & 1u) * TC_OUT_TILE;
                const int c0 = g0 + half * TC_CHUNK;
                const bool mine = c0 < p.BN && n0 + c0 < p.Cout;
                uint32_t acc[TC_CHUNK];
                if (mine) tmem_ld32(trow + (uint32_t)c0, acc);
                if (elected) bulk_wait_read<1>();
                epi_barrier();
                if (mine) {
                    const int n = n0 + c0;
                    const int ncols = min(TC_CHUNK, p.Cout - n);
                    const float* sbias = sBias + c0;
                    tmem_ld_wait(acc);
#pragma unroll
                    for (int h = 0; h < TC_CHUNK / 8; ++h) {
                        const float4 b0v = *reinterpret_cast<const float4*>(sbias + h * 8);
                        const float4 b1v = *reinterpret_cast<const float4*>(sbias + h * 8 + 4);
                        const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            v[e] = fminf(fmaxf(__uint_as_float(acc[h * 8 + e]) + bb[e], act_lo), act_hi);
                        if (p.res && row_ok && h * 8 < ncols) {
                            const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + row_off + n + h * 8));
                            const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(rh[e]);
                                v[2 * e] += f.x; v[2 * e + 1] += f.y;
                            }
                        }
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                        *reinterpret_cast<uint4*>(buf + r * 128 + (((half * 4 + h) ^ (r & 7)) << 4)) = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                epi_barrier();
                if (elected) {
                    if (p.mode4d) tma_store_4d(&map_o, smem_addr(buf), n0 + g0, ox0, oy0, b0);
                    else          tma_store_2d(&map_o, smem_addr(buf), n0 + g0, tm_ * TC_BM);
                    bulk_commit();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * a);     // the leader's MMA thread waits for both CTAs
        }
        if (elected) bulk_wait_all();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_barrier();                                           // nobody leaves while the peer may still signal or be read
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ======================================================================================
// Fused DepthwiseConv2D 3x3 -> 1x1 Conv2D (the tail of a MobileNetV2 inverted-residual block).
//
//   warp 0      TMA: per k-block (64 expanded channels) the INPUT PATCH of the tile -- (bw-1)s+3 x (bh-1)s+3 x bb
//               positions x 64 channels, one 4-D box, zero-filled outside the image (= the convolution padding) --
//               and the 1x1 weight tile B;
//   warps 12-19 depthwise: 3x3 taps from the swizzled patch in shared memory, + bias, activation, fp16, written
//               straight into the 128-byte-swizzled K-major A tile the UMMA descriptor expects;
//   warp 1      tcgen05.mma into double-buffered TMEM accumulators;
//   warps 4-11  epilogue: bias / activation / residual -> swizzled staging tile -> TMA store (as conv_tcgen05_kernel).
//
// The expanded activation is read from L2/HBM exactly once (by TMA, fully asynchronous, prefetched dw_pstages deep)
// and the depthwise output never exists in global memory.
__global__ void __launch_bounds__(TC_THREADS_DW, 1)
conv_dwproj_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                           const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o,
                           const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_stage = TC_BM * TC_BK * 2;                 // 16 KB
    const uint32_t b_stage = (uint32_t)p.BN * TC_BK * 2;
    unsigned char* sA = smem;
    unsigned char* sB = sA + TC_DW_ASTAGES * a_stage;
    unsigned char* sP = sB + TC_DW_ASTAGES * ((b_stage + 1023u) & ~1023u);
    unsigned char* sOut = sP + (size_t)p.dw_pstages * p.dw_patch_stage;
    float* sBias = reinterpret_cast<float*>(sOut + 2 * TC_OUT_TILE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + TC_OUT_BYTES);
    // bars: full[2] | empty[2] | pfull[4] | pempty[4] | tmem_full[2] | tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t b_pitch = (b_stage + 1023u) & ~1023u;

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_addr(bars), bar_empty = smem_addr(bars + 2);
    const uint32_t bar_pfull = smem_addr(bars + 4), bar_pempty = smem_addr(bars + 8);
    const uint32_t bar_tfull = smem_addr(bars + 12), bar_tempty = smem_addr(bars + 14);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_o);
        for (int s = 0; s < TC_DW_ASTAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1 + TC_DW_WARPS);        // B bytes (expect_tx) + one arrive per depthwise warp
            mbar_init(bar_empty + 8 * s, 1);                     // tcgen05.commit
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(bar_pfull + 8 * s, 1);                     // patch bytes (expect_tx)
            mbar_init(bar_pempty + 8 * s, TC_DW_WARPS);          // every depthwise warp is done with the patch
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_addr(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    const int per_img = p.tiles_w * p.tiles_h;

    // setmaxnreg at the top of each role's branch: the scheduler and epilogue warpgroups release registers, the depthwise
    // warpgroups (the stage that bounds the kernel) take them -- no spills inside the depthwise loop
    if (warp < TC_DW_EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_DW_REGS_SCHED));
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0, s = 0, ps = 0, tslot = 0;
            uint32_t ph = 0, pph = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const int tb = t / per_img, tr = t - tb * per_img;
                const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
                const int b0 = tb * p.bb, oy0 = th * p.bh, ox0 = tw * p.bw;
                for (int kb = 0; kb < p.n_kblocks; ++kb, ++it) {
                    if (it >= p.dw_pstages) mbar_wait(bar_pempty + 8 * ps, pph ^ 1u);
                    trace_stamp(p.trace, 0, tslot, 1);
                    mbar_expect_tx(bar_pfull + 8 * ps, p.dw_patch_bytes + 9 * 128);
                    tma_load_4d(smem_addr(sP + (size_t)ps * p.dw_patch_stage), &map_x, bar_pfull + 8 * ps, kb * TC_BK,
                                ox0 * p.dw_stride - p.dw_pad_l, oy0 * p.dw_stride - p.dw_pad_t, b0);
                    // the k-block's depthwise filter: 9 taps x 64 channels behind the patch (own 1024-byte aligned atom)
                    tma_load_2d(smem_addr(sP + (size_t)ps * p.dw_patch_stage + p.dw_patch_stage - 2048), &map_w, bar_pfull + 8 * ps,
                                kb * TC_BK, 0);
                    if (++ps == p.dw_pstages) { ps = 0; pph ^= 1u; }
                    if (it >= TC_DW_ASTAGES) mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    mbar_expect_tx(bar_full + 8 * s, p.b_bytes);
                    tma_load_2d(smem_addr(sB + (size_t)s * b_pitch), &map_b, bar_full + 8 * s, kb * TC_BK, 0);
                    if (++s == TC_DW_ASTAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int j = 0, s = 0, tslot = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++j) {
                const int a = j & 1;
                if (j >= 2) mbar_wait(bar_tempty + 8 * a, ((j >> 1) - 1) & 1);
                trace_stamp(p.trace, 1, tslot, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)a * p.acc_cols;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * s, ph);
                    trace_stamp(p.trace, 1, tslot, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc_sw128(smem_addr(sA + (size_t)s * a_stage));
                    const uint64_t db = umma_desc_sw128(smem_addr(sB + (size_t)s * b_pitch));
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        umma_f16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (kb > 0) || k > 0);
                    umma_commit(bar_empty + 8 * s);
                    if (++s == TC_DW_ASTAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(bar_tfull + 8 * a);
            }
        }
    }
    } else if (warp >= TC_DW_WARP0) {
        // ===================== depthwise 3x3 from the shared-memory patch =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_DW_REGS_DW));
        const int pt = (int)threadIdx.x - 32 * TC_DW_WARP0;
        // pair mode (<= 32 channels, stride 1, even box width: MobileNetV2's block 0): only the 4 live 16-byte chunks are
        // spread over the threads, each owns 2 horizontally adjacent pixels (12 patch loads and 72 HFMA2 instead of the
        // quad mode's 18 / 144 with half of the threads on dead channels)
        const int pairm = p.dw_pair;
        const int j = pairm ? (pt & 3) : (pt & 7), rg = pairm ? (pt >> 2) : (pt >> 3);      // 16-byte channel chunk, row group
        const int C8 = p.dw_C8, st = p.dw_stride, pw = p.dw_pw, php = p.dw_ph;
        const float lo = p.dw_act == SSD_ACT_NONE ? -__int_as_float(0x7f800000) : 0.0f;
        const float hi = p.dw_act == SSD_ACT_RELU6 ? 6.0f : __int_as_float(0x7f800000);
        const int box_rows = p.bw * p.bh * p.bb;
        // rows owned by this thread: strided (rg + 32 i) or, in quad mode, 4 horizontally adjacent pixels (4 rg + i)
        const int quad = p.dw_quad;
        int q0[TC_DW_ROWS];                                      // patch position of tap (0,0) per owned row; -1: padding row
#pragma unroll
        for (int i = 0; i < TC_DW_ROWS; ++i) {
            const int r = pairm ? (i < 2 ? 2 * rg + i : TC_BM) : quad ? 4 * rg + i : rg + 4 * TC_DW_WARPS * i;
            if (r < box_rows && r < TC_BM) {
                const int dx = r % p.bw, qq = r / p.bw, dy = qq % p.bh, db = qq / p.bh;
                q0[i] = dx * st + pw * (dy * st + php * db);
            } else {
                q0[i] = -1;
            }
        }
        // Arithmetic: packed half2 FMAs (4 per tap and row instead of 16 conversions + 8 FMAs).  The nine products of
        // an output are summed in fp16 -- the same storage format the result is rounded to for the tensor-core
        // operand -- starting from the fp16-rounded bias; the 1x1 projection accumulates in fp32 as everywhere else.
        int s = 0, ps = 0, it = 0, tslot = 0;
        uint32_t ph = 0, pph = 0;
        const bool has_bias = p.dw_bias != nullptr;
        if (pairm)       // logical chunks 4..7 of the A tiles are never written in pair mode: they must hold finite values
            for (int i = pt; i < TC_DW_ASTAGES * TC_BM * 4; i += 32 * TC_DW_WARPS) {      // (they meet zero-filled weights)
                const int stg = i / (TC_BM * 4), rr = (i >> 2) % TC_BM, c = 4 + (i & 3);
                *reinterpret_cast<uint4*>(sA + (size_t)stg * a_stage + rr * 128 + ((c ^ (rr & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
        auto load_bias = [&](int kb, __half2 (&b4)[4]) {
            const int c8 = kb * 8 + j;
            const bool okb = has_bias && c8 < C8;
            const float4 b0 = okb ? __ldg(reinterpret_cast<const float4*>(p.dw_bias) + c8 * 2) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = okb ? __ldg(reinterpret_cast<const float4*>(p.dw_bias) + c8 * 2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            b4[0] = __floats2half2_rn(b0.x, b0.y); b4[1] = __floats2half2_rn(b0.z, b0.w);
            b4[2] = __floats2half2_rn(b1.x, b1.y); b4[3] = __floats2half2_rn(b1.z, b1.w);
        };
        const __half2 lo2 = __float2half2_rn(p.dw_act == SSD_ACT_NONE ? -65504.0f : 0.0f);
        const __half2 hi2 = __float2half2_rn(p.dw_act == SSD_ACT_RELU6 ? 6.0f : 65504.0f);
        __half2 bias_next[4];
        load_bias(0, bias_next);
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            for (int kb = 0; kb < p.n_kblocks; ++kb, ++it) {
                const int c8 = kb * 8 + j;
                const bool cok = c8 < C8;
                __half2 acc[TC_DW_ROWS][4];
#pragma unroll
                for (int i = 0; i < TC_DW_ROWS; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[i][e] = bias_next[e];
                load_bias(kb + 1 < p.n_kblocks ? kb + 1 : 0, bias_next);              // in flight during this k-block
                mbar_wait(bar_pfull + 8 * ps, pph);                                   // patch + filter landed
                if (pt == 0) trace_stamp(p.trace, 2, tslot, 1);
                const unsigned char* patch = sP + (size_t)ps * p.dw_patch_stage;
                const unsigned char* wsm = patch + p.dw_patch_stage - 2048;
                if (pairm) {
                    // 2 outputs share 4 input columns per filter row
                    if (q0[0] >= 0) {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            uint4 w3[3];
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int tap = ky * 3 + kx;
                                w3[kx] = *reinterpret_cast<const uint4*>(wsm + tap * 128 + ((j ^ (tap & 7)) << 4));
                            }
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int q = q0[0] + c + pw * ky;
                                const uint4 xv = *reinterpret_cast<const uint4*>(patch + (size_t)q * 128 + ((j ^ (q & 7)) << 4));
                                const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                                for (int i = 0; i < 2; ++i) {
                                    const int kx = c - i;
                                    if (kx >= 0 && kx < 3) {
                                        const __half2* wh = reinterpret_cast<const __half2*>(&w3[kx]);
#pragma unroll
                                        for (int e = 0; e < 4; ++e) acc[i][e] = __hfma2(xh[e], wh[e], acc[i][e]);
                                    }
                                }
                            }
                        }
                    }
                } else if (quad) {
                    // sliding window: the 4 outputs of a thread share 6 input columns per filter row (18 loads, not 36)
                    if (q0[0] >= 0) {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            uint4 w3[3];
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int tap = ky * 3 + kx;
                                w3[kx] = *reinterpret_cast<const uint4*>(wsm + tap * 128 + ((j ^ (tap & 7)) << 4));
                            }
#pragma unroll
                            for (int c = 0; c < 6; ++c) {
                                const int q = q0[0] + c + pw * ky;
                                const uint4 xv = *reinterpret_cast<const uint4*>(patch + (size_t)q * 128 + ((j ^ (q & 7)) << 4));
                                const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int kx = c - i;
                                    if (kx >= 0 && kx < 3) {
                                        const __half2* wh = reinterpret_cast<const __half2*>(&w3[kx]);
#pragma unroll
                                        for (int e = 0; e < 4; ++e) acc[i][e] = __hfma2(xh[e], wh[e], acc[i][e]);
                                    }
                                }
                            }
                        }
                    }
                } else {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int tap = ky * 3 + kx;
                        const uint4 wvv = *reinterpret_cast<const uint4*>(wsm + tap * 128 + ((j ^ (tap & 7)) << 4));
                        const __half2* wh = reinterpret_cast<const __half2*>(&wvv);
#pragma unroll
                        for (int i = 0; i < TC_DW_ROWS; ++i) {
                            if (q0[i] < 0) continue;
                            const int q = q0[i] + kx + pw * ky;
                            const uint4 xv = *reinterpret_cast<const uint4*>(patch + (size_t)q * 128 + ((j ^ (q & 7)) << 4));
                            const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[i][e] = __hfma2(xh[e], wh[e], acc[i][e]);
                        }
                    }
                }
                if (pt == 0) trace_stamp(p.trace, 2, tslot, 2);
                if (it >= TC_DW_ASTAGES) mbar_wait(bar_empty + 8 * s, ph ^ 1u);         // A slot drained by the MMA
                if (pt == 0) trace_stamp(p.trace, 2, tslot, 3);
                unsigned char* a_tile = sA + (size_t)s * a_stage;
#pragma unroll
                for (int i = 0; i < TC_DW_ROWS; ++i) {
                    const int r = pairm ? (i < 2 ? 2 * rg + i : TC_BM) : quad ? 4 * rg + i : rg + 4 * TC_DW_WARPS * i;
                    if (r >= TC_BM) continue;
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (q0[i] >= 0 && cok) {
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) oh[e] = __hmin2(__hmax2(acc[i][e], lo2), hi2);
                    }
                    *reinterpret_cast<uint4*>(a_tile + r * 128 + ((j ^ (r & 7)) << 4)) = o;
                }
                if (pt == 0) trace_stamp(p.trace, 2, tslot, 5);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to the MMA
                if (pt == 0) trace_stamp(p.trace, 2, tslot, 6);
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_full + 8 * s); mbar_arrive(bar_pempty + 8 * ps); }
                if (pt == 0) trace_stamp(p.trace, 2, tslot, 4);
                if (++s == TC_DW_ASTAGES) { s = 0; ph ^= 1u; }
                if (++ps == p.dw_pstages) { ps = 0; pph ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (8 warps): TMEM -> bias/act/residual -> swizzled tile -> TMA store =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_DW_REGS_EPI));
        const int ew = warp - TC_DW_EPI_WARP0;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int r = q * 32 + lane;
        const bool elected = ew == 0 && lane == 0;
        const float act_lo = p.act == SSD_ACT_NONE ? -__int_as_float(0x7f800000) : 0.0f;
        const float act_hi = p.act == SSD_ACT_RELU6 ? 6.0f : __int_as_float(0x7f800000);
        {
            const int i = ew * 32 + lane;                        // bias of the (single) N tile, once
            sBias[i] = (p.bias && i < p.Cout) ? __ldg(p.bias + i) : 0.0f;
        }
        int j = 0, tslot = 0;
        uint32_t n_groups = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++j) {
            const int a = j & 1;
            if (elected) trace_stamp(p.trace, 3, tslot, 1);
            mbar_wait(bar_tfull + 8 * a, (j >> 1) & 1);
            if (elected) trace_stamp(p.trace, 3, tslot, 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int b = 0, pix = 0;
            const bool row_ok = tc_row_to_pixel(p, t, r, b, pix);
            const long long row_off = (long long)b * p.img0 + (long long)pix * p.pix0;
            const uint32_t trow = tmem_base + (uint32_t)a * p.acc_cols + ((uint32_t)(q * 32) << 16);
            const int tb = t / per_img, tr = t - tb * per_img;
            const int th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
            const int b0 = tb * p.bb, oy0 = th * p.bh, ox0 = tw * p.bw;
            for (int g0 = 0; g0 < p.BN && g0 < p.Cout; g0 += 64, ++n_groups) {
                unsigned char* buf = sOut + (n_groups & 1u) * TC_OUT_TILE;
                const int c0 = g0 + half * TC_CHUNK;
                const bool mine = c0 < p.BN && c0 < p.Cout;
                uint32_t acc[TC_CHUNK];
                if (mine) tmem_ld32(trow + (uint32_t)c0, acc);
                if (elected) bulk_wait_read<1>();
                epi_barrier();
                if (mine) {
                    const int ncols = min(TC_CHUNK, p.Cout - c0);
                    const float* sbias = sBias + c0;
                    tmem_ld_wait(acc);
#pragma unroll
                    for (int h = 0; h < TC_CHUNK / 8; ++h) {
                        const float4 b0v = *reinterpret_cast<const float4*>(sbias + h * 8);
                        const float4 b1v = *reinterpret_cast<const float4*>(sbias + h * 8 + 4);
                        const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            v[e] = fminf(fmaxf(__uint_as_float(acc[h * 8 + e]) + bb[e], act_lo), act_hi);
                        if (p.res && row_ok && h * 8 < ncols) {
                            const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + row_off + c0 + h * 8));
                            const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(rh[e]);
                                v[2 * e] += f.x; v[2 * e + 1] += f.y;
                            }
                        }
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                        *reinterpret_cast<uint4*>(buf + r * 128 + (((half * 4 + h) ^ (r & 7)) << 4)) = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                epi_barrier();
                if (elected) {
                    tma_store_4d(&map_o, smem_addr(buf), g0, ox0, oy0, b0);
                    bulk_commit();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * a);
            if (elected) trace_stamp(p.trace, 3, tslot, 3);
        }
        if (elected) bulk_wait_all();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// Deterministic split-K reduction + epilogue: partial[z][row][n] summed for z = 0..splits-1 in order.
__global__ void __launch_bounds__(256)
conv_splitk_reduce_kernel(const TcParams p, int rows_total) {
    pdl_trigger();
    pdl_wait();
    // one thread per (row, 4 consecutive channels): 16-byte loads of the partial planes, one row -> pixel
    // decomposition per 4 outputs
    const int n4 = (p.Cout + 3) >> 2;
    const int64_t total = (int64_t)rows_total * n4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / n4), n = (int)(e - (int64_t)row * n4) << 2;
        int b, pix;
        if (!tc_row_to_pixel(p, row / TC_BM, row % TC_BM, b, pix)) continue;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < p.splits; ++z) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(p.partial + ((size_t)z * rows_total + row) * p.ldp + n));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (n + k < p.Cout) tc_store_one(p, b, pix, n + k, vv[k]);
    }
}

// ------------------------------------------------------------------- host side --
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    }
    return fn;
}

static int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* elem_strides = nullptr, int swizzle_bytes = 128) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(SSD_ERR_UNSUPPORTED, "conv_tcgen05: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t bx[4], es[4];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SSD_ERR_UNSUPPORTED, "conv_tcgen05: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SSD_OK;
}

// Encoded tensor maps are cached per (pointer, geometry): cuTensorMapEncodeTiled costs a few
// microseconds of host time, which would otherwise dominate eager launches of small layers.
struct MapKey {
    const void* base; uint64_t dims[4]; uint64_t strides[3]; uint32_t box[4]; uint32_t es[4]; int rank; int swizzle;
    bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
        size_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
};
static std::mutex g_map_mutex;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

int cached_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.rank = rank; k.swizzle = swizzle_bytes;
    for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.es[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides_bytes[i];
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(k);
    if (it != g_map_cache.end()) { *map = it->second; return SSD_OK; }
    int rc = make_map(map, base, rank, dims, strides_bytes, box, elem_strides, swizzle_bytes);
    if (rc == SSD_OK) {
        if (g_map_cache.size() > 4096) g_map_cache.clear();
        g_map_cache.emplace(k, *map);
    }
    return rc;
}

// split-K workspaces: owned by the library (the only allocations it makes; they happen outside stream capture because
// plans are warmed up before capture).  Each split-K convolution gets its OWN region, keyed by (device, output pointer),
// carved from 64 MB arenas: launches of different layers may run concurrently on different streams (the multibox heads
// are parallel branches of the captured graph), and a region's address never changes once handed out, so CUDA graphs
// captured earlier stay valid.  Arenas are never freed.
struct PartialRegion { float* ptr; size_t bytes; };
struct PartialKey {
    int dev; const void* out;
    bool operator==(const PartialKey& o) const { return dev == o.dev && out == o.out; }
};
struct PartialKeyHash {
    size_t operator()(const PartialKey& k) const { return std::hash<const void*>()(k.out) ^ ((size_t)k.dev * 0x9E3779B97F4A7C15ull); }
};
static std::unordered_map<PartialKey, PartialRegion, PartialKeyHash> g_partial_regions;
static unsigned char* g_arena[16] = {nullptr};
static size_t g_arena_bytes[16] = {0}, g_arena_used[16] = {0};
static std::mutex g_partial_mutex;

static int partial_workspace(const void* out_key, size_t bytes, float** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return fail(SSD_ERR_UNSUPPORTED, "conv_tcgen05: device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_partial_mutex);
    const PartialKey key{dev, out_key};
    auto it = g_partial_regions.find(key);
    if (it != g_partial_regions.end() && it->second.bytes >= bytes) { *out = it->second.ptr; return SSD_OK; }
    bytes = (bytes + 255) & ~(size_t)255;
    if (g_arena_bytes[dev] - g_arena_used[dev] < bytes) {
        const size_t want = bytes < ((size_t)64 << 20) ? ((size_t)64 << 20) : bytes;
        void* fresh = nullptr;
        cudaError_t e = cudaMalloc(&fresh, want);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tcgen05: split-K workspace");
        g_arena[dev] = static_cast<unsigned char*>(fresh);        // the previous arena stays alive for captured graphs
        g_arena_bytes[dev] = want;
        g_arena_used[dev] = 0;
    }
    float* ptr = reinterpret_cast<float*>(g_arena[dev] + g_arena_used[dev]);
    g_arena_used[dev] += bytes;
    g_partial_regions[key] = PartialRegion{ptr, bytes};
    *out = ptr;
    return SSD_OK;
}

static int pair_mode_from_env() {
    const char* e = getenv("SSD_B200_PAIR");
    return e ? atoi(e) : -1;
}
static int g_pair_mode = pair_mode_from_env();
void conv_tcgen05_set_pair_mode(int mode) { g_pair_mode = mode; }

bool conv_tcgen05_supported(const ssd_conv_desc* d) {
    return (d->stride == 1 || d->stride == 2) && d->Cin % 8 == 0 && d->KH == d->KW && d->KH * d->KW <= 49 &&
           (reinterpret_cast<uintptr_t>(d->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0;
}

int conv_tcgen05_launch(const ssd_conv_desc* d, cudaStream_t st) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    const int taps = d->KH * d->KW;
    p.mode4d = !(taps == 1 && d->stride == 1 && d->pad_top == 0 && d->pad_left == 0 && d->Ho == d->H && d->Wo == d->W);
    p.stride = d->stride;
    p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo; p.HoWo = d->Ho * d->Wo; p.M = d->B * p.HoWo;
    p.Cin = d->Cin; p.KW = d->KW; p.dil = d->dilation; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
    p.kb_per_tap = (d->Cin + TC_BK - 1) / TC_BK;
    p.n_kblocks = taps * p.kb_per_tap;
    p.Cout = d->Cout;
    p.bias = d->bias; p.res = (const __half*)d->residual; p.out0 = d->out0; p.out1 = d->out1;
    p.act = d->act; p.out_f32 = d->out_f32; p.split = d->split;
    p.img0 = d->img_stride0; p.pix0 = d->pix_stride0; p.img1 = d->img_stride1; p.pix1 = d->pix_stride1;

    // N tile: the largest multiple of 16 (<= 256) that tiles Cout evenly; whole Cout when it fits one UMMA
    const int cout16 = (d->Cout + 15) / 16 * 16;
    // (with several N tiles the tile width is a multiple of 64: the TMA-store boxes are 64 channels wide and must
    // not reach into the neighbouring N tile; a ragged last tile is clipped by the tensor bounds)
    p.BN = 128;
    if (cout16 <= 256) p.BN = cout16;
    else for (int bn = 256; bn >= 128; bn -= 64) if (cout16 % bn == 0) { p.BN = bn; break; }
    int tiles_m;
    CUtensorMap map_a, map_b;
    if (!p.mode4d) {
        tiles_m = (p.M + TC_BM - 1) / TC_BM;
        uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)p.M};
        uint64_t str[1] = {(uint64_t)d->Cin * 2};
        uint32_t box[2] = {TC_BK, TC_BM};
        int rc = cached_map(&map_a, d->in, 2, dims, str, box);
        if (rc) return rc;
        p.a_bytes = TC_BM * TC_BK * 2;
    } else {
        // choose the output-pixel box (bw x bh x bb <= 128) that wastes the fewest MMA rows
        double best = -1.0;
        const int max_bw = d->stride == 1 ? TC_BM : 256 / d->stride;          // TMA box dimensions are limited to 256 elements
        for (int bw = 1; bw <= min(d->Wo, min(TC_BM, max_bw)); ++bw) {
            const int tw = (d->Wo + bw - 1) / bw;
            for (int bh = 1; bh <= min(d->Ho, min(TC_BM / bw, max_bw)); ++bh) {
                const int th = (d->Ho + bh - 1) / bh;
                int bb = 1;
                if (bw == d->Wo && bh == d->Ho) bb = max(1, min(d->B, TC_BM / (bw * bh)));
                const int tb = (d->B + bb - 1) / bb;
                const double eff = (double)p.M / ((double)tw * th * tb * TC_BM) + 1e-6 * bw;
                if (eff > best) { best = eff; p.bw = bw; p.bh = bh; p.bb = bb; p.tiles_w = tw; p.tiles_h = th; }
            }
        }
        tiles_m = p.tiles_w * p.tiles_h * ((d->B + p.bb - 1) / p.bb);
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        const uint32_t sst = (uint32_t)d->stride;
        uint32_t box[4] = {TC_BK, (uint32_t)p.bw * sst, (uint32_t)p.bh * sst, (uint32_t)p.bb};
        uint32_t est[4] = {1, sst, sst, 1};
        int rc = cached_map(&map_a, d->in, 4, dims, str, box, est);
        if (rc) return rc;
        p.a_bytes = (uint32_t)(p.bw * p.bh * p.bb) * TC_BK * 2;
    }
    // A layer whose tile grid leaves more than half of the SMs idle (M of a few thousand pixels: the 10x10 and 5x5 maps)
    // gets narrower N tiles instead of split-K: the same number of CTAs without the fp32 partial planes and without the
    // second launch that reduces them.  The narrowest of 128 / 64 that still fits one wave is taken.
    const int sms = sm_count();
    if (tiles_m * ((d->Cout + p.BN - 1) / p.BN) * 2 <= sms && p.BN >= 128 && d->Cout % 64 == 0) {
        for (int bn = 128; bn >= 64; bn >>= 1)
            if (bn < p.BN && d->Cout % bn == 0 && tiles_m * (d->Cout / bn) <= sms) p.BN = bn;
    }
    const int tiles_n = (d->Cout + p.BN - 1) / p.BN;
    p.acc_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
    p.tmem_cols = 2 * p.acc_cols;                                  // double-buffered accumulator (<= 512 columns)
    p.idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);   // f16 x f16 -> f32, K-major A and B
    {
        const uint64_t ktot = (uint64_t)taps * d->Cin;
        uint64_t dims[2] = {ktot, (uint64_t)d->Cout};
        uint64_t str[1] = {ktot * 2};
        uint32_t box[2] = {TC_BK, (uint32_t)p.BN};
        int rc = cached_map(&map_b, d->weight, 2, dims, str, box);
        if (rc) return rc;
        p.b_bytes = (uint32_t)p.BN * TC_BK * 2;
    }

    // split-K when the tile grid cannot fill the SMs and K is deep (the multibox head)
    int splits = 1;
    const int ctas = tiles_m * tiles_n;
    if (ctas * 4 <= sms && p.n_kblocks >= 8) {                     // (a grid that fills a quarter of the SMs runs as it is: the
                                                                   //  reduction launch costs more than the idle SMs)
        splits = min(min((sms + ctas - 1) / ctas, p.n_kblocks / 4), 32);
        if (splits < 1) splits = 1;
    }
    p.kb_per_split = (p.n_kblocks + splits - 1) / splits;
    splits = (p.n_kblocks + p.kb_per_split - 1) / p.kb_per_split;
    p.splits = splits;
    const int rows_total = tiles_m * TC_BM;
    if (splits > 1) {
        p.ldp = tiles_n * p.BN;
        int rc = partial_workspace(d->out0, (size_t)splits * rows_total * p.ldp * sizeof(float), &p.partial);
        if (rc) return rc;
    }

    p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.n_tiles = tiles_m * tiles_n * splits;

    // fp16 single-segment outputs leave through TMA stores (64-channel boxes of the NHWC output)
    CUtensorMap map_o;
    memset(&map_o, 0, sizeof(map_o));
    p.tma_store = 0;
    if (splits == 1 && !d->out_f32 && d->split >= d->Cout && (d->Cout & 7) == 0 && (p.pix0 & 7) == 0 && (p.img0 & 7) == 0 &&
        (reinterpret_cast<uintptr_t>(d->out0) & 15) == 0 && (d->residual == nullptr || (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0) &&
        (p.mode4d || p.img0 == (long long)p.HoWo * p.pix0)) {
        int rc;
        if (p.mode4d) {
            uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
            uint64_t str[3] = {(uint64_t)p.pix0 * 2, (uint64_t)d->Wo * p.pix0 * 2, (uint64_t)p.img0 * 2};
            uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bb};
            rc = cached_map(&map_o, d->out0, 4, dims, str, box);
        } else {
            uint64_t dims[2] = {(uint64_t)d->Cout, (uint64_t)p.M};
            uint64_t str[1] = {(uint64_t)p.pix0 * 2};
            uint32_t box[2] = {64, TC_BM};
            rc = cached_map(&map_o, d->out0, 2, dims, str, box);
        }
        if (rc) return rc;
        p.tma_store = 1;
    }
    // ---- CTA pair (cta_group::2, M = 256): deep-K layers with full 128 / 256-wide N tiles and at least one tile per SM
    //      (the VGG16 body and its gradients).  -1: automatic, 0: never, 1: whenever the shape allows (tests).
    if (splits == 1 && p.tma_store && (p.BN == 256 || p.BN == 128) && d->Cout % p.BN == 0 && p.n_kblocks >= 8 &&
        g_pair_mode != 0 && (g_pair_mode == 1 || tiles_m * tiles_n >= sms)) {
        CUtensorMap map_b2;
        {
            const uint64_t ktot = (uint64_t)taps * d->Cin;
            uint64_t dims[2] = {ktot, (uint64_t)d->Cout};
            uint64_t str[1] = {ktot * 2};
            uint32_t box[2] = {TC_BK, (uint32_t)p.BN / 2};
            int rc = cached_map(&map_b2, d->weight, 2, dims, str, box);
            if (rc) return rc;
        }
        TcParams q = p;
        q.b_bytes = (uint32_t)(p.BN / 2) * TC_BK * 2;
        q.idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        auto smem_pair = [&](int stages) {
            return (size_t)stages * (TC_BM * TC_BK * 2 + (size_t)(p.BN / 2) * TC_BK * 2) + (size_t)TC_OUT_BYTES +
                   (2 * TC_PAIR_STAGES + 4) * 8 + 16 + 1024;
        };
        q.stages = min(TC_PAIR_STAGES, p.n_kblocks);
        while (q.stages > 2 && smem_pair(q.stages) > (size_t)226 * 1024) --q.stages;
        static thread_local int pair_attr_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (pair_attr_dev != dev) {
            cudaError_t e = cudaFuncSetAttribute(conv_tcgen05_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return cuda_fail(e, "conv_tcgen05 (pair): cudaFuncSetAttribute");
            pair_attr_dev = dev;
        }
        const int n_pair_tiles = ((tiles_m + 1) / 2) * tiles_n;
        const int pairs = max(1, min(n_pair_tiles, sms / 2));
        conv_tcgen05_pair_kernel<<<dim3(2 * pairs), dim3(TC_THREADS), smem_pair(q.stages), st>>>(map_a, map_b2, map_o, q);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) return cuda_fail(le, "conv_tcgen05_pair_kernel");
        return SSD_OK;
    }

    // Operand ring depth and CTAs per SM.  A tile with few k-blocks (1x1 convolutions with Cin <= 128: every MobileNetV2
    // expand layer up to block 13) cannot use a deep ring; its persistent loop is bound by the latency of the epilogue's
    // dependent instruction chain, so a shallower ring that lets TWO CTAs share an SM (shared memory <= 113 KB and
    // TMEM <= 256 columns each) doubles the warps that hide that latency.
    auto smem_for = [&](int stages) {
        return (size_t)stages * (TC_BM * TC_BK * 2 + (size_t)p.BN * TC_BK * 2) + (size_t)TC_OUT_BYTES +
               (2 * TC_STAGES + 4) * 8 + 16 + 1024;
    };
    p.stages = min(TC_STAGES, max(2, p.kb_per_split));
    int ctas_per_sm = 1;
    if (p.n_tiles > sms && 2 * p.tmem_cols <= 512) {
        const int floor_stages = p.kb_per_split <= 2 ? 2 : 3;
        for (int st_ = p.stages; st_ >= floor_stages; --st_)
            if (2 * (smem_for(st_) + 1024) <= (size_t)227 * 1024) { p.stages = st_; ctas_per_sm = 2; break; }
    }
    const size_t smem = smem_for(p.stages);
    static thread_local int attr_dev = -1;                       // opt-in once per (thread, device): maximum footprint
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tcgen05: cudaFuncSetAttribute");
        attr_dev = cur_dev;
    }
    // persistent: one CTA per SM (512 TMEM columns and ~100-215 KB of shared memory per CTA)
    dim3 grid(min(p.n_tiles, sms * ctas_per_sm), 1, 1);
    {
        cudaError_t le = launch_pdl(conv_tcgen05_kernel, grid, dim3(TC_THREADS), smem, st, map_a, map_b, map_o, p);
        if (le != cudaSuccess) return cuda_fail(le, "conv_tcgen05_kernel");
    }
    if (splits > 1) {
        const int64_t total = (int64_t)rows_total * ((p.Cout + 3) / 4);
        const int64_t want = (total + 255) / 256, cap = (int64_t)sms * 8;
        int blocks = (int)(want < cap ? want : cap);
        cudaError_t le = launch_pdl(conv_splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, p, rows_total);
        if (le != cudaSuccess) return cuda_fail(le, "conv_splitk_reduce_kernel");
    }
    return SSD_OK;
}

}  // namespace ssd
// Debug / test hook: -1 automatic (CTA pairs for layers with at least one tile per SM), 0 never, 1 whenever the shape allows.
extern "C" int ssd_debug_pair_mode(int mode) {
    ssd::conv_tcgen05_set_pair_mode(mode);
    return SSD_OK;
}
namespace ssd {

// Fused DepthwiseConv2D 3x3 (+ folded BN + activation) -> 1x1 Conv2D (+ folded BN, residual): the "depthwise ->
// project" tail of a MobileNetV2 inverted-residual block as ONE launch.  GEMM view: M = B*Ho*Wo pixels,
// K = C (expanded channels, computed on the fly by the depthwise warps), N = Cout <= 256.
// Tile / ring geometry of the fused kernel; returns false when no output box fits the shared-memory budget.
static bool dwproj_plan(const ssd_dwproj_desc* d, TcParams* pp, size_t* smem_out) {
    TcParams& p = *pp;
    p.BN = (d->Cout + 15) / 16 * 16;
    const size_t b_pitch = ((size_t)p.BN * TC_BK * 2 + 1023) & ~(size_t)1023;
    const size_t fixed = (size_t)TC_DW_ASTAGES * (TC_BM * TC_BK * 2 + b_pitch) + (size_t)TC_OUT_BYTES + 17 * 8 + 16 + 1024;
    const size_t budget = (size_t)227 * 1024;
    const int M = d->B * d->Ho * d->Wo, s = d->stride;
    double best = -1.0;
    for (int bw = 1; bw <= min(d->Wo, TC_BM); ++bw) {
        const int tw = (d->Wo + bw - 1) / bw;
        for (int bh = 1; bh <= min(d->Ho, TC_BM / bw); ++bh) {
            const int th = (d->Ho + bh - 1) / bh;
            int bb = 1;
            if (bw == d->Wo && bh == d->Ho) bb = max(1, min(d->B, TC_BM / (bw * bh)));
            const int tb = (d->B + bb - 1) / bb;
            const int pw = (bw - 1) * s + 3, ph = (bh - 1) * s + 3;
            if (pw > 256 || ph > 256) continue;
            const size_t patch = (((size_t)pw * ph * bb * 128 + 1023) & ~(size_t)1023) + 2048;   // + the k-block's filter
            int pst = 0;
            for (int c = 3; c >= 2; --c) if (fixed + c * patch <= budget) { pst = c; break; }
            if (!pst) continue;
            // MMA-row efficiency, discounted by the halo the patch re-reads; stride-1 boxes whose width is a multiple of 4
            // run the sliding-window depthwise (half the shared-memory loads)
            const double eff = (double)M / ((double)tw * th * tb * TC_BM) * ((double)bw * bh * s * s / ((double)pw * ph)) *
                               ((s == 1 && bw % 4 == 0) ? 1.25 : 1.0) + 1e-6 * bw;
            if (eff > best) {
                best = eff; p.bw = bw; p.bh = bh; p.bb = bb; p.tiles_w = tw; p.tiles_h = th;
                p.dw_pw = pw; p.dw_ph = ph; p.dw_patch_bytes = (uint32_t)(pw * ph * bb * 128); p.dw_patch_stage = (uint32_t)patch;
                p.dw_pstages = pst;
                p.n_tiles = tw * th * tb;
                *smem_out = fixed + pst * patch;
            }
        }
    }
    return best > 0.0;
}

bool conv_dwproj_supported(const ssd_dwproj_desc* d) {
    if (!(d->C % 8 == 0 && d->Cout % 8 == 0 && d->Cout <= 256 && (d->stride == 1 || d->stride == 2) &&
          (reinterpret_cast<uintptr_t>(d->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->dw_weight) & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(d->proj_weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
          (d->residual == nullptr || (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0) &&
          (d->dw_bias == nullptr || (reinterpret_cast<uintptr_t>(d->dw_bias) & 15) == 0)))
        return false;
    TcParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0;
    return dwproj_plan(d, &p, &smem);
}

int conv_dwproj_launch(const ssd_dwproj_desc* d, cudaStream_t st) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0;
    if (!dwproj_plan(d, &p, &smem)) return fail(SSD_ERR_UNSUPPORTED, "ssd_dwproj: no tile geometry fits shared memory");
    p.trace = debug_trace_buffer();
    if (p.trace)
        fprintf(stderr, "ssd_dwproj: box %dx%dx%d patch %dx%d tiles=%d kblocks=%d pstages=%d quad=%d BN=%d smem=%zu\n", p.bw, p.bh,
                p.bb, p.dw_pw, p.dw_ph, p.n_tiles, p.n_kblocks, p.dw_pstages, p.dw_quad, p.BN, smem);
    p.mode4d = 1;
    p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo; p.HoWo = d->Ho * d->Wo; p.M = d->B * p.HoWo;
    p.Cin = d->C; p.KW = 1; p.dil = 1; p.stride = 1;
    p.kb_per_tap = (d->C + TC_BK - 1) / TC_BK;
    p.n_kblocks = p.kb_per_tap;
    p.kb_per_split = p.n_kblocks;
    p.Cout = d->Cout;
    p.bias = d->proj_bias; p.res = (const __half*)d->residual; p.out0 = d->out; p.out1 = nullptr;
    p.act = d->act; p.out_f32 = 0; p.split = d->Cout;
    p.pix0 = d->Cout; p.img0 = (long long)p.HoWo * d->Cout;
    p.acc_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
    p.tmem_cols = 2 * p.acc_cols;
    p.idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    p.dw_x = reinterpret_cast<const uint4*>(d->in); p.dw_w = reinterpret_cast<const uint4*>(d->dw_weight); p.dw_bias = d->dw_bias;
    p.dw_H = d->H; p.dw_W = d->W; p.dw_C8 = d->C / 8; p.dw_stride = d->stride; p.dw_pad_t = d->pad_top; p.dw_pad_l = d->pad_left;
    p.dw_act = d->dw_act;
    p.splits = 1; p.tiles_m = p.n_tiles; p.tiles_n = 1; p.tma_store = 1; p.stages = TC_DW_ASTAGES;
    p.dw_quad = (d->stride == 1 && p.bw % 4 == 0) ? 1 : 0;
    p.dw_pair = (d->stride == 1 && p.bw % 2 == 0 && d->C <= 32) ? 1 : 0;

    CUtensorMap map_x, map_w, map_b, map_o;
    {
        uint64_t dims[2] = {(uint64_t)d->C, 9};
        uint64_t str[1] = {(uint64_t)d->C * 2};
        uint32_t box[2] = {TC_BK, 9};
        int rc = cached_map(&map_w, d->dw_weight, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->C * 2, (uint64_t)d->W * d->C * 2, (uint64_t)d->H * d->W * d->C * 2};
        uint32_t box[4] = {TC_BK, (uint32_t)p.dw_pw, (uint32_t)p.dw_ph, (uint32_t)p.bb};
        int rc = cached_map(&map_x, d->in, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->C, (uint64_t)d->Cout};
        uint64_t str[1] = {(uint64_t)d->C * 2};
        uint32_t box[2] = {TC_BK, (uint32_t)p.BN};
        int rc = cached_map(&map_b, d->proj_weight, 2, dims, str, box);
        if (rc) return rc;
        p.b_bytes = (uint32_t)p.BN * TC_BK * 2;
    }
    {
        uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)d->Wo * d->Cout * 2, (uint64_t)p.HoWo * d->Cout * 2};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bb};
        int rc = cached_map(&map_o, d->out, 4, dims, str, box);
        if (rc) return rc;
    }
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaError_t e = cudaFuncSetAttribute(conv_dwproj_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "conv_dwproj: cudaFuncSetAttribute");
        attr_dev = cur_dev;
    }
    dim3 grid(min(p.n_tiles, sm_count()), 1, 1);
    cudaError_t le = launch_pdl(conv_dwproj_tcgen05_kernel, grid, dim3(TC_THREADS_DW), smem, st, map_x, map_w, map_b, map_o, p);
    if (le != cudaSuccess) return cuda_fail(le, "conv_dwproj_tcgen05_kernel");
    return SSD_OK;
}

// ======================================================================================
// Filter gradient on tcgen05:  dW[co, tap, ci] += sum_pixels dY[pix, co] * X[pix (+) tap, ci]
//
// Both operands are pixel-major in memory (NHWC), i.e. "MN-major" for the tensor core: the
// reduction dimension K (pixels) is the ROW index of both shared-memory tiles.  A pixel tile is a
// (bw x bh) box of one image; TMA loads dY[box, 64 co] and X[box shifted by the tap, 64 ci] as
// [64 pixels][128 B] swizzled tiles (out-of-image pixels are zero-filled, which is exactly the
// convolution padding), one tile per 64-channel group.  UMMA descriptors: MN-major, SWIZZLE_128B,
// SBO = 1024 B (next 8 pixel rows), LBO = 64*128 B (next 64-channel group), +2048 B per UMMA_K = 16.
// Accumulator 128 (co) x BN (ci) fp32 in TMEM; split-K over pixel tiles across blockIdx.z,
// combined with vector fp32 atomics.
constexpr int WG5_KP = 64;            // pixels per k-block
constexpr int WG5_STAGES = 4;
constexpr int WG5_THREADS = 192;      // warp 0 TMA, warp 1 MMA/TMEM, warps 2-5 epilogue

struct Wg5Params {
    int Cout, Cin, taps, KW, dil, pad_t, pad_l;
    int bw, bh, tiles_w, tiles_h, B;          // pixel tiles: (tiles_w * tiles_h) per image
    int n_ptiles, ptiles_per_split;
    int BN, n_groups;                         // ci tile (multiple of 64), groups of 64
    int tiles_n;
    uint32_t idesc, tmem_cols, a_bytes, b_bytes;
    float* dw;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(WG5_THREADS, 1)
conv_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                          const Wg5Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t grp = WG5_KP * 128;                           // one [64 px][128 B] tile = 8 KB
    const uint32_t a_stage = 2 * grp, b_stage = (uint32_t)p.n_groups * grp;
    unsigned char* sA = smem;
    unsigned char* sB = smem + WG5_STAGES * a_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + WG5_STAGES * b_stage);     // full[S] | empty[S] | done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG5_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_addr(bars), bar_empty = smem_addr(bars + WG5_STAGES), bar_done = smem_addr(bars + 2 * WG5_STAGES);

    const int tile_m = blockIdx.x / p.tiles_n, tile_n = blockIdx.x - tile_m * p.tiles_n;
    const int co0 = tile_m * 128, ci0 = tile_n * p.BN;
    const int tap = blockIdx.y, ky = tap / p.KW, kx = tap - ky * p.KW;
    const int pt0 = blockIdx.z * p.ptiles_per_split, pt1 = min(p.n_ptiles, pt0 + p.ptiles_per_split);
    const int nkb = pt1 - pt0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < WG5_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_addr(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int per_img = p.tiles_w * p.tiles_h;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % WG5_STAGES;
                if (i >= WG5_STAGES) mbar_wait(bar_empty + 8 * s, ((i / WG5_STAGES) - 1) & 1);
                const int pt = pt0 + i;
                const int b = pt / per_img, tr = pt - b * per_img, th = tr / p.tiles_w, tw = tr - th * p.tiles_w;
                const int oy0 = th * p.bh, ox0 = tw * p.bw;
                mbar_expect_tx(bar_full + 8 * s, p.a_bytes + p.b_bytes);
                const uint32_t dst_a = smem_addr(sA + (size_t)s * a_stage), dst_b = smem_addr(sB + (size_t)s * b_stage);
                tma_load_4d(dst_a, &map_dy, bar_full + 8 * s, co0, ox0, oy0, b);
                tma_load_4d(dst_a + grp, &map_dy, bar_full + 8 * s, co0 + 64, ox0, oy0, b);
                for (int g = 0; g < p.n_groups; ++g)
                    tma_load_4d(dst_b + g * grp, &map_x, bar_full + 8 * s, ci0 + g * 64, ox0 + kx * p.dil - p.pad_l,
                                oy0 + ky * p.dil - p.pad_t, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % WG5_STAGES;
                mbar_wait(bar_full + 8 * s, (i / WG5_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = umma_desc_mn_sw128(smem_addr(sA + (size_t)s * a_stage), grp);
                const uint64_t db = umma_desc_mn_sw128(smem_addr(sB + (size_t)s * b_stage), grp);
#pragma unroll
                for (int k = 0; k < WG5_KP / 16; ++k)            // 16 pixel rows = 2048 B per UMMA_K
                    umma_f16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), p.idesc, (i | k) != 0);
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_done);
        }
    } else {
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            if (ci0 + c0 >= p.Cin) break;
            uint32_t acc[16];
            tmem_ld16(trow + (uint32_t)c0, acc);
            if (co < p.Cout) {
                float* dst = p.dw + ((size_t)co * p.taps + tap) * p.Cin + ci0 + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    if (ci0 + c0 + j < p.Cin)
                        atomicAdd(reinterpret_cast<float4*>(dst + j),
                                  make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                              __uint_as_float(acc[j + 3])));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

bool conv_wgrad_tcgen05_supported(const ssd_conv_desc* d, int ldy) {
    return d->stride == 1 && d->Cin % 64 == 0 && d->KH == d->KW && ldy % 8 == 0 &&
           (reinterpret_cast<uintptr_t>(d->in) & 15) == 0;
}

int conv_wgrad_tcgen05_launch(const ssd_conv_desc* d, const void* d_dy, int ldy, float* d_dw, cudaStream_t st) {
    Wg5Params p;
    memset(&p, 0, sizeof(p));
    p.Cout = d->Cout; p.Cin = d->Cin; p.taps = d->KH * d->KW; p.KW = d->KW; p.dil = d->dilation;
    p.pad_t = d->pad_top; p.pad_l = d->pad_left; p.B = d->B; p.dw = d_dw;
    // pixel box of 64 output pixels with the least out-of-image waste
    double best = -1.0;
    for (int bw = 1; bw <= WG5_KP; bw *= 2) {
        const int bh = WG5_KP / bw;
        const int tw = (d->Wo + bw - 1) / bw, th = (d->Ho + bh - 1) / bh;
        const double eff = (double)d->Wo * d->Ho / ((double)tw * bw * th * bh) + 1e-6 * bw;
        if (eff > best) { best = eff; p.bw = bw; p.bh = bh; p.tiles_w = tw; p.tiles_h = th; }
    }
    p.n_ptiles = p.tiles_w * p.tiles_h * d->B;
    p.BN = d->Cin >= 256 ? 256 : d->Cin;                       // Cin is a multiple of 64
    p.n_groups = p.BN / 64;
    p.tiles_n = (d->Cin + p.BN - 1) / p.BN;
    const int tiles_m = (d->Cout + 127) / 128;
    p.tmem_cols = p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
    p.idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // MN-major A and B
    p.a_bytes = 2 * WG5_KP * 128;
    p.b_bytes = (uint32_t)p.n_groups * WG5_KP * 128;

    CUtensorMap map_dy, map_x;
    {
        uint64_t dims[4] = {(uint64_t)ldy, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)ldy * 2, (uint64_t)d->Wo * ldy * 2, (uint64_t)d->Ho * d->Wo * ldy * 2};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
        int rc = cached_map(&map_dy, d_dy, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
        int rc = cached_map(&map_x, d->in, 4, dims, str, box);
        if (rc) return rc;
    }
    const int sms = sm_count();
    const int base = tiles_m * p.tiles_n * p.taps;
    int splits = max(1, min((2 * sms + base - 1) / base, (p.n_ptiles + 7) / 8));
    p.ptiles_per_split = (p.n_ptiles + splits - 1) / splits;
    splits = (p.n_ptiles + p.ptiles_per_split - 1) / p.ptiles_per_split;
    const size_t smem = (size_t)WG5_STAGES * (p.a_bytes + p.b_bytes) + (2 * WG5_STAGES + 1) * 8 + 16 + 1024;
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "conv_wgrad_tcgen05: cudaFuncSetAttribute");
        attr_dev = cur_dev;
    }
    dim3 grid(tiles_m * p.tiles_n, p.taps, splits);
    conv_wgrad_tcgen05_kernel<<<grid, WG5_THREADS, smem, st>>>(map_dy, map_x, p);
    SSD_CHECK_LAUNCH("conv_wgrad_tcgen05_kernel");
    return SSD_OK;
}

}  // namespace ssd
