// Shared helpers for libssd_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/ssd_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libssd_b200 is written for sm_100a (B200) only"
#endif

namespace ssd {

// thread-local last-error text, exposed through ssd_last_error()
char* last_error_buffer();
int   fail(int code, const char* fmt, ...);
int   cuda_fail(cudaError_t e, const char* what);

#define SSD_REQUIRE_PTR(p)                                                   \
    do { if ((p) == nullptr) return ::ssd::fail(SSD_ERR_NULL, "%s: %s is NULL", __func__, #p); } while (0)
#define SSD_REQUIRE(cond, code, ...)                                         \
    do { if (!(cond)) return ::ssd::fail((code), __VA_ARGS__); } while (0)
#define SSD_CHECK_LAUNCH(what)                                               \
    do { cudaError_t e__ = cudaGetLastError();                               \
         if (e__ != cudaSuccess) return ::ssd::cuda_fail(e__, what); } while (0)

static inline cudaStream_t as_stream(ssd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();      // cached multiProcessorCount of the current device

// Debug tracing (ssd_debug_trace, tools/trace_kernel.py): when a device buffer is registered, CTA 0 of the instrumented
// kernels records (globaltimer << 8 | tag) stamps, 8 roles x 512 slots; nullptr (the default) compiles to one
// predictable branch per stamp.
unsigned long long* debug_trace_buffer();
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int role, int& slot, int tag) {
    if (trace && blockIdx.x == 0 && blockIdx.y == 0 && slot < 512) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        trace[role * 512 + slot] = (t << 8) | (unsigned long long)(tag & 0xff);
        ++slot;
    }
}

// ---- programmatic dependent launch (PDL) -------------------------------------
// Consecutive kernels of a plan are launched with programmatic stream serialisation:
// kernel N+1 may start its prologue (barrier init, TMEM allocation, descriptor
// prefetch, shared-memory weight staging) while kernel N drains.  Every kernel
// launched through launch_pdl() MUST call pdl_wait() before it touches global memory
// written or read by its predecessors; pdl_trigger() (at kernel entry) lets the
// successor be scheduled as soon as all CTAs of this grid have started.
// On by default (SSD_B200_PDL=0 switches it off; off automatically under Nsight Compute / compute-sanitizer): see
// pdl_enabled() in common.cu.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void pdl_wait()    { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int    ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device-side float32 arithmetic with exactly one IEEE rounding per op ----
// The box kernels must reproduce a chain of separate TensorFlow ops bit for
// bit, so products are never contracted into FMAs (the TU is also built with
// -fmad=false; the intrinsics make the intent explicit and robust).
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte asynchronous global -> shared copy (LDGSTS, bypasses L1 and the register file): a
// thread can have many of these in flight, so a whole staging tile is requested before the
// first byte arrives instead of paying one DRAM round trip per loop iteration.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Cooperative copy of `total` consecutive floats between global memory and a shared
// staging region with 16-byte global accesses AND 16-byte shared accesses: the data is
// placed in shared memory at the same misalignment (mod 4 floats) as in global memory,
// so both sides of every vector move are aligned.  `region` must be 16-byte aligned and
// hold total + 3 floats.  Returns the pointer p with p[e] == src[e].
// The vector part is issued with cp.async: call stage_rows_wait() (then __syncthreads())
// before reading the rows; several regions may be issued back to back before one wait.
__device__ __forceinline__ float* stage_rows_in(const float* __restrict__ src, int total, float* __restrict__ region) {
    const int mis = (int)(((uintptr_t)src >> 2) & 3);          // floats past a 16-byte boundary
    float* s = region + mis;
    const int head = min(total, (4 - mis) & 3);
    const int nvec = (total - head) >> 2;
    const float4* v = reinterpret_cast<const float4*>(src + head);
    float4* d = reinterpret_cast<float4*>(s + head);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) cp_async16(d + i, v + i);
    for (int e = threadIdx.x; e < head; e += blockDim.x) s[e] = __ldcs(src + e);
    for (int e = head + (nvec << 2) + threadIdx.x; e < total; e += blockDim.x) s[e] = __ldcs(src + e);
    return s;
}
__device__ __forceinline__ void stage_rows_wait() { cp_async_wait_all(); }
// The reverse: s must have been laid out with the destination's misalignment, i.e.
// s == region + ((uintptr_t)dst >> 2 & 3).
__device__ __forceinline__ float* stage_rows_ptr(const float* dst, float* region) {
    return region + (int)(((uintptr_t)dst >> 2) & 3);
}
__device__ __forceinline__ void stage_rows_out(float* __restrict__ dst, int total, const float* __restrict__ s) {
    const int mis = (int)(((uintptr_t)dst >> 2) & 3);
    const int head = min(total, (4 - mis) & 3);
    const int nvec = (total - head) >> 2;
    for (int e = threadIdx.x; e < head; e += blockDim.x) __stcs(dst + e, s[e]);
    float4* v = reinterpret_cast<float4*>(dst + head);
    const float4* d = reinterpret_cast<const float4*>(s + head);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) __stcs(v + i, d[i]);
    for (int e = head + (nvec << 2) + threadIdx.x; e < total; e += blockDim.x) __stcs(dst + e, s[e]);
}

}  // namespace ssd
