// Micro-benchmark: register-tile Rot application (the gate kernel's inner loop) on sm_100a.
// Variants: matrix from shared memory (3-register FFMA), matrix from __constant__ memory with a
// warp-uniform index (uniform-register / constant-bank FFMA operand), and the RZ-RY-RZ decomposition
// (merged diagonal phases + real rotations).  Reports executed FP32 instruction rate per variant.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int RB = 5, R = 32, NG = 640;   // gates per sweep (wire = g % RB)
__constant__ float c_gates[NG * 8];

struct Mat { float r00, i00, r01, i01, r10, i10, r11, i11; };
__device__ __forceinline__ void apply_pair(const Mat &m, float2 &x0, float2 &x1) {
    const float2 a = x0, b = x1;
    x0.x = m.r00 * a.x - m.i00 * a.y + m.r01 * b.x - m.i01 * b.y;
    x0.y = m.r00 * a.y + m.i00 * a.x + m.r01 * b.y + m.i01 * b.x;
    x1.x = m.r10 * a.x - m.i10 * a.y + m.r11 * b.x - m.i11 * b.y;
    x1.y = m.r10 * a.y + m.i10 * a.x + m.r11 * b.y + m.i11 * b.x;
}
template <int Q> __device__ __forceinline__ void gate_q(const Mat &m, float2 (&s)[R]) {
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
        const int r0 = ((j >> Q) << (Q + 1)) | (j & ((1 << Q) - 1));
        apply_pair(m, s[r0], s[r0 | (1 << Q)]);
    }
}
template <int Q> __device__ __forceinline__ void ry_q(float c, float sn, float2 (&s)[R]) {
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
        const int r0 = ((j >> Q) << (Q + 1)) | (j & ((1 << Q) - 1));
        const int r1 = r0 | (1 << Q);
        const float2 a = s[r0], b = s[r1];
        s[r0].x = c * a.x - sn * b.x;  s[r0].y = c * a.y - sn * b.y;
        s[r1].x = sn * a.x + c * b.x;  s[r1].y = sn * a.y + c * b.y;
    }
}

// MODE 0: smem matrices; 1: constant memory; 2: decomposed (smem tables)
template <int MODE> __global__ void __launch_bounds__(128) k_rot(const float *gates, float *out, int sweeps, const int *zeros) {
    __shared__ float4 gs4[NG * 2];
    float *gs = reinterpret_cast<float *>(gs4);
    for (int i = threadIdx.x; i < NG * 8; i += blockDim.x) gs[i] = gates[i];
    __syncthreads();
    const int nu_t = MODE == 3 ? zeros[threadIdx.x] : 0;
    float2 s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = make_float2(1e-3f * (threadIdx.x + r), 1e-3f * r);
    for (int sw = 0; sw < sweeps; ++sw) {
#pragma unroll 1
        for (int g = 0; g < NG; g += RB) {
            if (MODE == 2) {
                // diag (pre), 5 real rotations, diag (post): tables of 32 complex phases each, broadcast reads
                const float2 *tab = reinterpret_cast<const float2 *>(gs) + (g % 64) * 8;
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const float2 p = tab[r], a = s[r];
                    s[r] = make_float2(a.x * p.x - a.y * p.y, a.x * p.y + a.y * p.x);
                }
                const float2 cs0 = tab[32], cs1 = tab[33], cs2 = tab[34], cs3 = tab[35], cs4 = tab[36];
                ry_q<0>(cs0.x, cs0.y, s); ry_q<1>(cs1.x, cs1.y, s); ry_q<2>(cs2.x, cs2.y, s);
                ry_q<3>(cs3.x, cs3.y, s); ry_q<4>(cs4.x, cs4.y, s);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const float2 p = tab[40 + r], a = s[r];
                    s[r] = make_float2(a.x * p.x - a.y * p.y, a.x * p.y + a.y * p.x);
                }
            } else {
                Mat m[RB];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    if (MODE == 0 || MODE == 3) {
                        const int nu = MODE == 3 ? nu_t : 0;   // MODE 3: address not provably uniform
                        const float4 a = gs4[(g + q + nu) * 2], b = gs4[(g + q + nu) * 2 + 1];
                        m[q] = Mat{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    } else {
                        const float *c = c_gates + (g + q) * 8;
                        m[q] = Mat{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]};
                    }
                }
                gate_q<0>(m[0], s); gate_q<1>(m[1], s); gate_q<2>(m[2], s); gate_q<3>(m[3], s); gate_q<4>(m[4], s);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) acc += s[r].x + s[r].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3;
}

int main() {
    float h[NG * 8];
    for (int i = 0; i < NG * 8; ++i) h[i] = 0.35f * ((i * 2654435761u >> 8) % 1000) / 1000.f - 0.17f + ((i % 8 == 0 || i % 8 == 6) ? 0.9f : 0.f);
    float *dg, *out;
    cudaMalloc(&dg, sizeof(h)); cudaMemcpy(dg, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_gates, h, sizeof(h));
    cudaMalloc(&out, 148 * 16 * 128 * 4);
    int *dz; cudaMalloc(&dz, 128 * 4); cudaMemset(dz, 0, 128 * 4);
    const int sweeps = 20;
    for (int cps = 1; cps <= 4; ++cps) {       // CTAs per SM (128 threads each): 4/8/12/16 warps per SM
        const int grid = 148 * cps;
        const double gates = (double)grid * 128 * sweeps * NG;      // thread-gates
        const double instr_full = gates * 16 * 16, instr_dec = gates / RB * (2 * 31 * 4 + 5 * 16 * 8);
        float t0 = timeit([&] { k_rot<0><<<grid, 128>>>(dg, out, sweeps, dz); });
        float t1 = timeit([&] { k_rot<1><<<grid, 128>>>(dg, out, sweeps, dz); });
        float t2 = timeit([&] { k_rot<2><<<grid, 128>>>(dg, out, sweeps, dz); });
        float t3 = timeit([&] { k_rot<3><<<grid, 128>>>(dg, out, sweeps, dz); });
        printf("CTAs/SM %d: 3-register FFMA (matrix in vector registers) %.3f ms %.1f Tinstr-flop/s\n", cps, t3, instr_full * 2 / t3 / 1e9);
        printf("CTAs/SM %d: smem-mat %.3f ms %.1f Tinstr-flop/s | const-mat %.3f ms %.1f | decomposed %.3f ms %.1f (speedup vs smem-mat %.2fx)\n",
               cps, t0, instr_full * 2 / t0 / 1e9, t1, instr_full * 2 / t1 / 1e9, t2, instr_dec * 2 / t2 / 1e9, t0 / t2);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
