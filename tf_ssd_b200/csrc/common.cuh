// Shared helpers for libssd_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

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

}  // namespace ssd
