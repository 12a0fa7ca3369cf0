// Linear layers with one NARROW side next to the circuits: `linear_down` (pixels -> qubits) and `linear_up`
// (qubits -> pixels) of the re-upload families (nn/qdense.py:565-670, :219-386; torch.nn.Linear in float64 there).
// y = x W^T + b with min(in, out) <= 16 is not a GEMM worth a tensor core: it streams the wide tensor once and does
// <= 16 FMAs per element, so it belongs at the HBM roofline -- the library's float64 GEMMs ran these shapes at a
// third of it and a separate reduction computed the bias gradient.  Three streaming kernels cover forward and backward
// of both orientations (W is addressed through strides, so the transposed products of the backward reuse them):
//   narrow_out_kernel : y[r, j]  = sum_k x[r, k] W(j, k) (+ b[j])          one warp per row, J accumulators per lane
//   wide_out_kernel   : y[r, n]  = sum_j h[r, j] W(n, j) (+ b[n])          a thread owns its n's, W(n, :) in registers
//   outer_kernel      : P[c][n][j] = sum_{r in chunk c} a[r, j] b[r, n],  Pb[c][n] = sum b[r, n]   (then a fixed-order sum
//                       over the row chunks: deterministic, no float atomics)
// Arithmetic in the tensors' dtype (float64 like the reference modules, or float32).
#include <cuda_runtime.h>
#include <stdint.h>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

constexpr int MAXJ = 16;       // narrow side
constexpr int NPT = 4;         // wide-side elements a thread owns (256 threads -> 1024 per CTA column chunk)
constexpr int OPT = 4;         // outer_kernel: owned columns per thread
constexpr int ORW = 4;         // outer_kernel: rows in flight per thread and column (2 x 8 with three CTAs per SM measured slower:
                               // 74 vs 58 us on the 40960 x 784 float64 shapes)
constexpr int SUB = 64;        // rows of the narrow tensor staged in shared memory at a time

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[r, j] = sum_k x[r, k] * W[j * sj + k * sk] + bias[j];  one warp per row.  W is
// staged once per CTA in shared memory as [j][k] (consecutive lanes -> consecutive words, whatever its global strides are;
// SMEM = false: read through L1 instead, for weights that do not fit).
template <typename T, int J, bool SMEM>
__global__ void __launch_bounds__(256) narrow_out_kernel(const T *x, const T *W, long long sj, long long sk, const T *bias, T *y,
                                                         long long rows, int K) {
    extern __shared__ __align__(16) unsigned char lin_smem[];
    T *ws = reinterpret_cast<T *>(lin_smem);
    if (SMEM) {
        for (int i = threadIdx.x; i < J * K; i += blockDim.x) {
            const int jj = i / K, k = i - jj * K;
            ws[i] = W[jj * sj + k * sk];
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    // one row per warp; eight strides (8 independent loads per lane) per trip.  (16-byte loads were tried: a lane stride of 16
    // bytes halves the shared-memory wavefront efficiency of the weight reads -- 22 M wavefronts per launch, slower.)
    for (long long r = warp0; r < rows; r += nwarps) {
        const T *xr = x + r * K;
        T acc[J];
#pragma unroll
        for (int jj = 0; jj < J; ++jj) acc[jj] = (T)0;
        int k = lane;
        for (; k + 224 < K; k += 256) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(xr + k + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int jj = 0; jj < J; ++jj)
                    acc[jj] += v[u] * (SMEM ? ws[jj * K + k + 32 * u] : __ldg(W + jj * sj + (k + 32 * u) * sk));
        }
        for (; k < K; k += 32) {
            const T v0 = __ldg(xr + k);
#pragma unroll
            for (int jj = 0; jj < J; ++jj) acc[jj] += v0 * (SMEM ? ws[jj * K + k] : __ldg(W + jj * sj + k * sk));
        }
#pragma unroll
        for (int jj = 0; jj < J; ++jj) acc[jj] = warp_sum(acc[jj]);
        if (lane == 0) {
#pragma unroll
            for (int jj = 0; jj < J; ++jj) y[r * J + jj] = acc[jj] + (bias != nullptr ? bias[jj] : (T)0);
        }
    }
}

// float64, J <= 8: the same product on the FP64 tensor cores (mma.sync.m8n8k4.f64): a warp takes 8 rows, the weights sit in
// shared memory as [8][K + 4] (rows >= J zero; the pad keeps the 8 x 4 fragment reads at the two-wavefront minimum), per
// k-step one 8-byte load per lane (8 rows x 32 contiguous bytes), one shared-memory read, one DMMA.  The scalar kernel above
// spends J shared-memory reads + J DFMAs per element and is bound by shared-memory wavefronts (ncu: 22 M per launch, 80 of
// its 120 us); here the weights are read once per 8 rows x 4 columns.
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256, 4) narrow_out_dmma_kernel(const double *x, const double *W, long long sj, long long sk,
                                                              const double *bias, double *y, long long rows, int K, int J) {
    extern __shared__ __align__(16) unsigned char lin_smem[];
    double *ws = reinterpret_cast<double *>(lin_smem);
    const int ldw = K + 4;
    for (int i = threadIdx.x; i < 8 * ldw; i += blockDim.x) {
        const int jj = i / ldw, k = i - jj * ldw;
        ws[i] = (jj < J && k < K) ? W[jj * sj + k * sk] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const double *wrow = ws + g * ldw + tig;
    for (long long r0 = 8 * warp0; r0 < rows; r0 += 8 * nwarps) {
        const long long r = r0 + g < rows ? r0 + g : rows - 1;          // rows past the end re-read the last one (never stored)
        const double *xr = x + r * K + tig;
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;                  // two accumulator pairs: independent DMMA chains
        int k0 = 0;
        // 16 k-steps per trip: 16 loads (4 KB per warp) in flight per lane, four CTAs per SM -- the pass is bound by bytes in
        // flight (Little: ~2 us loaded latency x 6.5 TB/s / 148 SMs = ~90 KB per SM; 32 KB ran at 2.4 TB/s)
        for (; k0 + 64 <= K; k0 += 64) {
            double a[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) a[u] = __ldg(xr + k0 + 4 * u);
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
                dmma_m8n8k4(c0, c1, a[u], wrow[k0 + 4 * u]);
                dmma_m8n8k4(d0, d1, a[u + 1], wrow[k0 + 4 * u + 4]);
            }
        }
        for (; k0 < K; k0 += 4) {
            const double av = k0 + tig < K ? __ldg(xr + k0) : 0.0;      // the zero pad of the weights covers k >= K as well
            dmma_m8n8k4(c0, c1, av, wrow[k0]);
        }
        c0 += d0; c1 += d1;
        if (r0 + g < rows) {
            const int n = 2 * tig;
            if (n < J) y[(r0 + g) * J + n] = c0 + (bias != nullptr ? bias[n] : 0.0);
            if (n + 1 < J) y[(r0 + g) * J + n + 1] = c1 + (bias != nullptr ? bias[n + 1] : 0.0);
        }
    }
}

// y[r, n] = sum_j h[r, j] * W[n * sn + j * sj] + bias[n];  grid (row chunks, n chunks), thread owns n = n0 + tid + i * 256
template <typename T, int J>
__global__ void __launch_bounds__(256) wide_out_kernel(const T *h, const T *W, long long sn, long long sj, const T *bias, T *y,
                                                       long long rows, int N, long long rows_per_chunk) {
    const int n0 = blockIdx.y * (256 * NPT);
    T w[NPT][J], bv[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
        const int n = n0 + threadIdx.x + i * 256;
        bv[i] = (n < N && bias != nullptr) ? bias[n] : (T)0;
#pragma unroll
        for (int j = 0; j < J; ++j) w[i][j] = n < N ? W[n * sn + j * sj] : (T)0;
    }
    const long long r0 = (long long)blockIdx.x * rows_per_chunk;
    const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    __shared__ T hs[SUB * J];                // the narrow rows of a sub-block: one coalesced load instead of a dependent
                                             // broadcast load per row
    for (long long rb = r0; rb < r1; rb += SUB) {
        const int nr = (int)(r1 - rb < SUB ? r1 - rb : SUB);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * J; i += 256) hs[i] = __ldg(h + rb * J + i);
        __syncthreads();
        for (int rr = 0; rr < nr; ++rr) {
            T hv[J];
#pragma unroll
            for (int j = 0; j < J; ++j) hv[j] = hs[rr * J + j];
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                const int n = n0 + threadIdx.x + i * 256;
                if (n < N) {
                    T a = bv[i];
#pragma unroll
                    for (int j = 0; j < J; ++j) a += hv[j] * w[i][j];
                    y[(rb + rr) * N + n] = a;
                }
            }
        }
    }
}

// P[c][n][j] = sum_{r in row chunk c} a[r, j] * b[r, n];  Pb[c][n] = sum_r b[r, n] (wide-side bias gradient) and
// Pa[c][j] = sum_r a[r, j] (narrow-side bias gradient; n chunk 0 only).  grid (row chunks, n chunks)
template <typename T, int J>
__global__ void __launch_bounds__(256) outer_kernel(const T *a, const T *b, T *P, T *Pb, T *Pa, long long rows, int N,
                                                    long long rows_per_chunk) {
    const int n0 = blockIdx.y * (256 * OPT);
    T acc[OPT][J], accb[OPT], acca = (T)0;
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        accb[i] = (T)0;
#pragma unroll
        for (int j = 0; j < J; ++j) acc[i][j] = (T)0;
    }
    const long long r0 = (long long)blockIdx.x * rows_per_chunk;
    const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    __shared__ T as[SUB * J];
    for (long long rb = r0; rb < r1; rb += SUB) {
        const int nr = (int)(r1 - rb < SUB ? r1 - rb : SUB);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * J; i += 256) as[i] = __ldg(a + rb * J + i);
        __syncthreads();
        if (Pa != nullptr && blockIdx.y == 0 && threadIdx.x < J)
            for (int rr = 0; rr < nr; ++rr) acca += as[rr * J + threadIdx.x];
        for (int rr = 0; rr < nr; rr += ORW) {        // ORW rows of the wide tensor in flight per thread and owned column
            T bvv[ORW][OPT];
#pragma unroll
            for (int u = 0; u < ORW; ++u)
#pragma unroll
                for (int i = 0; i < OPT; ++i) {
                    const int n = n0 + threadIdx.x + i * 256;
                    bvv[u][i] = (rr + u < nr && n < N) ? __ldg(b + (rb + rr + u) * N + n) : (T)0;
                }
#pragma unroll
            for (int u = 0; u < ORW; ++u) {
                T av[J];
#pragma unroll
                for (int j = 0; j < J; ++j) av[j] = rr + u < nr ? as[(rr + u) * J + j] : (T)0;
#pragma unroll
                for (int i = 0; i < OPT; ++i) {
                    accb[i] += bvv[u][i];
#pragma unroll
                    for (int j = 0; j < J; ++j) acc[i][j] += av[j] * bvv[u][i];
                }
            }
        }
    }
    const long long c = blockIdx.x;
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int n = n0 + threadIdx.x + i * 256;
        if (n < N) {
#pragma unroll
            for (int j = 0; j < J; ++j) P[(c * N + n) * J + j] = acc[i][j];
            if (Pb != nullptr) Pb[c * N + n] = accb[i];
        }
    }
    if (Pa != nullptr && blockIdx.y == 0 && threadIdx.x < J) Pa[c * J + threadIdx.x] = acca;
}

// out[i * so_n + j * so_j] = sum_c P[c][i][j]  (i < N, j < J), in a fixed order;  vectors: J == 1.  A CTA takes 32 elements;
// its 8 warps split the chunks (coalesced over the elements), the 8 partial sums are added in sequence.
template <typename T>
__global__ void __launch_bounds__(256) chunk_sum_kernel(const T *P, int chunks, long long n_elems, int J, long long so_n, long long so_j,
                                                        T *out) {
    __shared__ T part[8][33];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const long long e = (long long)blockIdx.x * 32 + lane;
    T s = (T)0;
    if (e < n_elems)
        for (int c = wp; c < chunks; c += 8) s += P[(long long)c * n_elems + e];
    part[wp][lane] = s;
    __syncthreads();
    if (wp == 0 && e < n_elems) {
        T t = (T)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][lane];
        const long long n = e / J, j = e - n * J;
        out[n * so_n + j * so_j] = t;
    }
}

struct Chunks { int row_chunks, n_chunks; long long rows_per_chunk; };
Chunks pick_chunks(long long rows, int N, int npt = NPT, int ctas_per_sm = 2) {
    Chunks c;
    c.n_chunks = (N + 256 * npt - 1) / (256 * npt);
    long long rc = (ctas_per_sm * 148 + c.n_chunks - 1) / c.n_chunks;
    if (rc > (rows + 31) / 32) rc = (rows + 31) / 32;                 // at least 32 rows per chunk
    if (rc < 1) rc = 1;
    c.rows_per_chunk = (rows + rc - 1) / rc;
    c.row_chunks = (int)((rows + c.rows_per_chunk - 1) / c.rows_per_chunk);
    return c;
}

#define QIDDM_DISPATCH_J(J_, ...)                                                                                         \
    switch (J_) {                                                                                                         \
        case 1: { constexpr int J = 1; __VA_ARGS__; break; }   case 2: { constexpr int J = 2; __VA_ARGS__; break; }                     \
        case 3: { constexpr int J = 3; __VA_ARGS__; break; }   case 4: { constexpr int J = 4; __VA_ARGS__; break; }                     \
        case 5: { constexpr int J = 5; __VA_ARGS__; break; }   case 6: { constexpr int J = 6; __VA_ARGS__; break; }                     \
        case 7: { constexpr int J = 7; __VA_ARGS__; break; }   case 8: { constexpr int J = 8; __VA_ARGS__; break; }                     \
        case 9: { constexpr int J = 9; __VA_ARGS__; break; }   case 10: { constexpr int J = 10; __VA_ARGS__; break; }                   \
        case 11: { constexpr int J = 11; __VA_ARGS__; break; } case 12: { constexpr int J = 12; __VA_ARGS__; break; }                   \
        case 13: { constexpr int J = 13; __VA_ARGS__; break; } case 14: { constexpr int J = 14; __VA_ARGS__; break; }                   \
        case 15: { constexpr int J = 15; __VA_ARGS__; break; } default: { constexpr int J = 16; __VA_ARGS__; break; }                   \
    }

template <typename T>
void launch_narrow_out(const T *x, const T *W, long long sj, long long sk, const T *bias, T *y, long long rows, int K, int Jn,
                       cudaStream_t s) {
    if (sizeof(T) == 8 && Jn <= 8 && (size_t)8 * (K + 4) * 8 <= 200 * 1024) {
        const size_t smem8 = (size_t)8 * (K + 4) * 8;
        const long long blocks8 = (rows + 63) / 64;                  // 8 warps x 8 rows
        const unsigned grid = (unsigned)(blocks8 < 148 * 4 ? blocks8 : 148 * 4);
        if (smem8 > 48 * 1024)
            cudaFuncSetAttribute(narrow_out_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        narrow_out_dmma_kernel<<<grid, 256, smem8, s>>>(reinterpret_cast<const double *>(x), reinterpret_cast<const double *>(W), sj, sk,
                                                        reinterpret_cast<const double *>(bias), reinterpret_cast<double *>(y), rows, K, Jn);
        count_launch();
        return;
    }
    const long long blocks = (rows + 7) / 8;                         // 8 warps, one row each per trip
    const size_t smem = (size_t)Jn * K * sizeof(T);
    if (smem <= 96 * 1024) {
        // persistent CTAs (the weights are staged once per CTA): as many per SM as their shared memory allows, up to four
        int per_sm = smem > 0 ? (int)((200 * 1024) / (smem + 1024)) : 4;
        per_sm = per_sm > 4 ? 4 : (per_sm < 1 ? 1 : per_sm);
        const unsigned grid = (unsigned)(blocks < 148 * per_sm ? blocks : 148 * per_sm);
        // above 48 KB the opt-in limit is needed (per device and function; setting it is a host-side call of a few microseconds)
        QIDDM_DISPATCH_J(Jn, {
            if (smem > 48 * 1024)
                cudaFuncSetAttribute(narrow_out_kernel<T, J, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            narrow_out_kernel<T, J, true><<<grid, 256, smem, s>>>(x, W, sj, sk, bias, y, rows, K);
        })
    } else {
        const unsigned grid = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
        QIDDM_DISPATCH_J(Jn, (narrow_out_kernel<T, J, false><<<grid, 256, 0, s>>>(x, W, sj, sk, bias, y, rows, K)))
    }
    count_launch();
}
template <typename T>
void launch_wide_out(const T *h, const T *W, long long sn, long long sj, const T *bias, T *y, long long rows, int N, int Jn,
                     cudaStream_t s) {
    const Chunks c = pick_chunks(rows, N);
    const dim3 grid(c.row_chunks, c.n_chunks);
    QIDDM_DISPATCH_J(Jn, (wide_out_kernel<T, J><<<grid, 256, 0, s>>>(h, W, sn, sj, bias, y, rows, N, c.rows_per_chunk)))
    count_launch();
}
// gW(n, j) (strides so_n, so_j) = sum_r a[r, j] b[r, n];  gb_wide[n] = sum_r b[r, n];  gb_narrow[j] = sum_r a[r, j]
template <typename T>
void launch_outer(const T *a, const T *b, T *gW, long long so_n, long long so_j, T *gb_wide, T *gb_narrow, long long rows, int N,
                  int Jn, T *ws, cudaStream_t s) {
    const Chunks c = pick_chunks(rows, N, OPT, 2);
    T *P = ws;
    T *Pb = gb_wide != nullptr ? P + (size_t)c.row_chunks * N * Jn : nullptr;
    T *Pa = gb_narrow != nullptr ? P + (size_t)c.row_chunks * N * (Jn + 1) : nullptr;
    const dim3 grid(c.row_chunks, c.n_chunks);
    QIDDM_DISPATCH_J(Jn, (outer_kernel<T, J><<<grid, 256, 0, s>>>(a, b, P, Pb, Pa, rows, N, c.rows_per_chunk)))
    const long long ne = (long long)N * Jn;
    chunk_sum_kernel<T><<<(unsigned)((ne + 31) / 32), 256, 0, s>>>(P, c.row_chunks, ne, Jn, so_n, so_j, gW);
    count_launch(2);
    if (gb_wide != nullptr) {
        chunk_sum_kernel<T><<<(unsigned)((N + 31) / 32), 256, 0, s>>>(Pb, c.row_chunks, N, 1, 1, 0, gb_wide);
        count_launch();
    }
    if (gb_narrow != nullptr) {
        chunk_sum_kernel<T><<<1, 256, 0, s>>>(Pa, c.row_chunks, Jn, 1, 1, 0, gb_narrow);
        count_launch();
    }
}

template <typename T>
int fwd_t(const void *x_, const void *w_, const void *b_, void *y_, long long rows, int in_f, int out_f, cudaStream_t s) {
    const T *x = reinterpret_cast<const T *>(x_), *W = reinterpret_cast<const T *>(w_), *b = reinterpret_cast<const T *>(b_);
    T *y = reinterpret_cast<T *>(y_);
    if (out_f <= in_f) launch_narrow_out<T>(x, W, in_f, 1, b, y, rows, in_f, out_f, s);       // W (out, in): j = out, k = in
    else launch_wide_out<T>(x, W, in_f, 1, b, y, rows, out_f, in_f, s);                       // n = out, j = in
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

template <typename T>
int bwd_t(const void *x_, const void *w_, const void *gy_, void *gx_, void *gw_, void *gb_, long long rows, int in_f, int out_f,
          void *ws, cudaStream_t s) {
    const T *x = reinterpret_cast<const T *>(x_), *W = reinterpret_cast<const T *>(w_), *gy = reinterpret_cast<const T *>(gy_);
    T *gx = reinterpret_cast<T *>(gx_), *gw = reinterpret_cast<T *>(gw_), *gb = reinterpret_cast<T *>(gb_);
    if (out_f <= in_f) {
        // narrow out (J = out, wide = in):  gx[r, k] = sum_j gy[r, j] W[j, k]  (wide-out product with W^T);
        // gW[j, k] = sum_r gy[r, j] x[r, k];  gb[j] = sum_r gy[r, j]
        if (gx != nullptr) launch_wide_out<T>(gy, W, 1, in_f, nullptr, gx, rows, in_f, out_f, s);
        if (gw != nullptr || gb != nullptr) {
            if (gw == nullptr) return QIDDM_EINVAL;
            launch_outer<T>(gy, x, gw, 1, in_f, nullptr, gb, rows, in_f, out_f, reinterpret_cast<T *>(ws), s);
        }
    } else {
        // wide out (J = in, wide = out):  gx[r, j] = sum_n gy[r, n] W[n, j]  (narrow-out product with W^T);
        // gW[n, j] = sum_r gy[r, n] x[r, j];  gb[n] = sum_r gy[r, n]
        if (gx != nullptr) launch_narrow_out<T>(gy, W, 1, in_f, nullptr, gx, rows, out_f, in_f, s);
        if (gw != nullptr || gb != nullptr) {
            if (gw == nullptr) return QIDDM_EINVAL;
            launch_outer<T>(x, gy, gw, in_f, 1, gb, nullptr, rows, out_f, in_f, reinterpret_cast<T *>(ws), s);
        }
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

}  // namespace

size_t skinny_linear_ws_bytes(long long rows, int in_f, int out_f) {
    const int wide = in_f > out_f ? in_f : out_f, narrow = in_f > out_f ? out_f : in_f;
    const Chunks c = pick_chunks(rows > 0 ? rows : 1, wide, OPT, 2);
    return (size_t)c.row_chunks * ((size_t)wide * (narrow + 1) + narrow) * 8 + 256;
}

int skinny_linear_forward(const void *x, const void *w, const void *bias, void *y, int dtype, long long rows, int in_f, int out_f,
                          cudaStream_t s) {
    if (!x || !w || !y || rows < 0 || in_f < 1 || out_f < 1) return QIDDM_EINVAL;
    if ((in_f < out_f ? in_f : out_f) > MAXJ) return QIDDM_EUNSUPPORTED;
    if (rows == 0) return QIDDM_OK;
    if (dtype == QIDDM_DTYPE_F64) return fwd_t<double>(x, w, bias, y, rows, in_f, out_f, s);
    if (dtype == QIDDM_DTYPE_F32) return fwd_t<float>(x, w, bias, y, rows, in_f, out_f, s);
    return QIDDM_EINVAL;
}

int skinny_linear_backward(const void *x, const void *w, const void *grad_y, void *grad_x, void *grad_w, void *grad_b, int dtype,
                           long long rows, int in_f, int out_f, void *ws, cudaStream_t s) {
    if (!x || !w || !grad_y || !ws || rows < 1 || in_f < 1 || out_f < 1) return QIDDM_EINVAL;
    if ((in_f < out_f ? in_f : out_f) > MAXJ) return QIDDM_EUNSUPPORTED;
    if (dtype == QIDDM_DTYPE_F64) return bwd_t<double>(x, w, grad_y, grad_x, grad_w, grad_b, rows, in_f, out_f, ws, s);
    if (dtype == QIDDM_DTYPE_F32) return bwd_t<float>(x, w, grad_y, grad_x, grad_w, grad_b, rows, in_f, out_f, ws, s);
    return QIDDM_EINVAL;
}

}  // namespace qiddm
