// Micro-benchmark: latency / throughput of legacy mma.sync.m16n8k16 (HMMA.16816.F32) on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/hmma_latency.cu -o /tmp/hmma && /tmp/hmma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(long long* out, float* sink, int iters) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a = 0x3c003c00u + threadIdx.x, b = 0x3c003c00u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) mma(c[i], a, a, a, a, b, b);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps) {
    long long* d; float* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4 * 1024 * 148);
    const int iters = 1000;
    k<CHAINS><<<1, 32 * warps>>>(d, sink, iters);
    k<CHAINS><<<1, 32 * warps>>>(d, sink, iters);
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("chains/warp=%d warps=%2d: %.1f cycles per MMA-round (%.1f cycles per MMA per SM)\n", CHAINS, warps, (double)h / iters,
           (double)h / iters / (CHAINS * warps));
    cudaFree(d); cudaFree(sink);
}

int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(4); run<2>(4); run<4>(4); run<1>(16); run<2>(16); run<4>(16);
    return 0;
}
