// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on sm_100a.
// Each thread runs N_CHAINS independent dependency chains; complex-rotation-like register pattern.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int CH> __global__ void k_scalar(float *out, float a, float b) {
    float x[CH], y[CH];
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = i * 0.5f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) { x[i] = fmaf(a, x[i], y[i]); y[i] = fmaf(b, y[i], x[i]); }
    }
    float s = 0; for (int i = 0; i < CH; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> __global__ void k_packed(float *out, float a, float b) {
    float2 x[CH], y[CH];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < CH; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i); y[i] = make_float2(i * 0.5f, 1.f); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) { x[i] = __ffma2_rn(a2, x[i], y[i]); y[i] = __ffma2_rn(b2, y[i], x[i]); }
    }
    float s = 0; for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y + y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int grid = 148 * 8, block = 256; constexpr int CH = 8;
    float ms1 = timeit([&] { k_scalar<CH><<<grid, block>>>(out, 0.999f, 1.001f); });
    float ms2 = timeit([&] { k_packed<CH><<<grid, block>>>(out, 0.999f, 1.001f); });
    double fl1 = (double)grid * block * ITERS * CH * 2 * 2, fl2 = fl1 * 2;
    printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms1, fl1 / ms1 / 1e9);
    printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", ms2, fl2 / ms2 / 1e9);
    return 0;
}
