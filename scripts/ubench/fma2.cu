// Micro-benchmark: FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a (diagnostic tooling).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float2& d, float2 a, float2 b) {
    unsigned long long da, aa, bb;
    aa = ((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x);
    bb = ((unsigned long long)__float_as_uint(b.y) << 32) | __float_as_uint(b.x);
    da = ((unsigned long long)__float_as_uint(d.y) << 32) | __float_as_uint(d.x);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(aa), "l"(bb));
    d.x = __uint_as_float((unsigned)da); d.y = __uint_as_float((unsigned)(da >> 32));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a0, float b0, int iters) {
    float2 acc[8];
    float2 a = make_float2(a0, a0 * 1.01f), b = make_float2(b0, b0 * 0.99f);
    for (int i = 0; i < 8; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
                else if (MODE == 1) { float2 t = acc[i]; fma2(t, a, b); acc[i].x = t.x * 1.0f; acc[i].y = t.y; acc[i] = t; }
                else { // mixed: packed fma + one ALU-pipe op (FMNMX) per packed fma
                    float2 t = acc[i]; fma2(t, a, b); t.x = fmaxf(t.x, -1e30f); acc[i] = t; }
            }
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fmas = 148.0 * 8 * 256 * iters * 4 * 8 * 2;   // scalar FMAs
            printf("mode %d: %.3f ms  %.2f TFLOP/s (2 flop/FMA)  err=%s\n", mode, ms, 2 * fmas / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
