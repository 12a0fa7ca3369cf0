// Density-matrix path for the MID-CIRCUIT noise channels of the re-upload classes (reference: the `add_noise` branches of
// nn/qdense.py:515-527 (differN_noise_befor), :1405-1417 (QIDDM_PL_noise), :1504-1516, :1599-1617 (QIDDM_LL_noise),
// :1693-1705, evaluated on PennyLane's `default.mixed` by src/mnist_noise.py:211-229 -- a trained net whose
// `add_noise` flag is flipped at test time; inference only).
//
// Per re-upload block the reference applies, on every wire j, RZ(a_j) followed by a single-qubit channel, then the
// StronglyEntanglingLayers block.  rho is kept TRANSPOSED, tau = rho^T (row c of tau = column c of rho, contiguous), and
// every step is expressed with pieces that already exist or stream well:
//   * channels.  PhaseDamping / AmplitudeDamping / DepolarizingChannel on one wire act on the 2 x 2 block of that wire as
//     (rho00, rho11) -> M (rho00, rho11) with a real column-stochastic M, and rho01, rho10 -> f_off * (rho01, rho10).  The
//     elements {(r, r ^ x)} of a fixed x = r XOR c are closed under all wires' channels, and on such a class the n channels
//     are the Kronecker product  (x_j ? f_off : M)  over the wires: one CTA per (instance, class) runs n butterfly stages
//     in shared memory.  RZ(a_j) is diagonal and commutes with every channel (it leaves the mixed diagonal pair alone and
//     multiplies each scaled off-diagonal element by a phase), so it moves behind the channels into the unitary part;
//   * unitary part  rho -> U rho U^dagger, U = SEL(W_i) RZ(a): the existing gate kernel applied to the 2^n rows of tau as
//     independent state vectors (QIDDM_INIT_STATE, the rows of one instance share its angles) gives (U rho)^T; its
//     conjugate transpose is conj(U rho); the same launch on that gives conj(U rho) U^T = conj(U rho U^dagger) = tau'
//     (rho' is Hermitian).  One gate launch pair and one transpose per block; no second transpose.
// Readout from the diagonal: probabilities (first K, scale, clamp) or <Z_j>.
#include <cuda_runtime.h>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

// tau[b][r][r ^ x] for all r: n butterfly stages (real 2 x 2 M on the wires where x has a 0 bit), f_off^popc(x) overall
__global__ void __launch_bounds__(256) dm_channel_kernel(float2 *tau, int n, float f_off, float m00, float m01, float m10,
                                                         float m11) {
    extern __shared__ float2 dm_sm[];
    const int A = 1 << n;
    const int x = blockIdx.x;
    float2 *base = tau + (size_t)blockIdx.y * A * A;
    float scale = 1.f;
    for (int q = 0; q < n; ++q)
        if ((x >> q) & 1) scale *= f_off;
    for (int r = threadIdx.x; r < A; r += blockDim.x) {
        const float2 v = base[(size_t)r * A + (r ^ x)];
        dm_sm[r] = make_float2(v.x * scale, v.y * scale);
    }
    __syncthreads();
    for (int q = 0; q < n; ++q) {
        if ((x >> q) & 1) continue;                    // block-uniform
        for (int i = threadIdx.x; i < A / 2; i += blockDim.x) {
            const int lo = i & ((1 << q) - 1);
            const int r0 = ((i >> q) << (q + 1)) | lo, r1 = r0 | (1 << q);
            const float2 a = dm_sm[r0], c = dm_sm[r1];
            dm_sm[r0] = make_float2(m00 * a.x + m01 * c.x, m00 * a.y + m01 * c.y);
            dm_sm[r1] = make_float2(m10 * a.x + m11 * c.x, m10 * a.y + m11 * c.y);
        }
        __syncthreads();
    }
    for (int r = threadIdx.x; r < A; r += blockDim.x) base[(size_t)r * A + (r ^ x)] = dm_sm[r];
}

// dst[b][i][j] = conj(src[b][j][i]), 32 x 32 tiles through shared memory (A >= 2)
__global__ void __launch_bounds__(256) dm_transpose_conj_kernel(const float2 *src, float2 *dst, int A) {
    __shared__ float2 tile[32][33];
    const float2 *s = src + (size_t)blockIdx.z * A * A;
    float2 *d = dst + (size_t)blockIdx.z * A * A;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int k = ty; k < 32; k += 8) {
        const int row = j0 + k, col = i0 + tx;
        if (row < A && col < A) tile[k][tx] = s[(size_t)row * A + col];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int row = i0 + k, col = j0 + tx;
        if (row < A && col < A) {
            const float2 v = tile[tx][k];
            d[(size_t)row * A + col] = make_float2(v.x, -v.y);
        }
    }
}

__global__ void __launch_bounds__(256) dm_init_kernel(float2 *tau, long long total, long long AA) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        tau[i] = make_float2((i % AA) == 0 ? 1.f : 0.f, 0.f);
}

// one CTA per instance: probabilities (read_count, stride, scale, clamp) or <Z_j> from the diagonal of tau
__global__ void __launch_bounds__(256) dm_readout_kernel(const float2 *tau, int n, int readout, int read_count, int read_stride,
                                                         float post_scale, int clamp, float lo, float hi, float *out) {
    __shared__ float red[8];
    const int A = 1 << n;
    const float2 *t = tau + (size_t)blockIdx.x * A * A;
    if (readout == QIDDM_READ_PROBS) {
        for (int m = threadIdx.x; m < read_count; m += blockDim.x) {
            const int k = m * read_stride;
            float v = post_scale * t[(size_t)k * A + k].x;
            if (clamp) v = fminf(fmaxf(v, lo), hi);
            out[(size_t)blockIdx.x * read_count + m] = v;
        }
        return;
    }
    for (int j = 0; j < n; ++j) {
        float acc = 0.f;
        for (int k = threadIdx.x; k < A; k += blockDim.x) {
            const float pr = t[(size_t)k * A + k].x;
            acc += ((k >> (n - 1 - j)) & 1) ? -pr : pr;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            out[(size_t)blockIdx.x * n + j] = post_scale * s;
        }
        __syncthreads();
    }
}

}  // namespace

size_t dm_state_bytes(int n_qubits, long long B) {
    return (((size_t)B << (2 * n_qubits)) * sizeof(float2) + 255) & ~(size_t)255;
}

int dm_init(float2 *tau, int n, long long B, cudaStream_t s) {
    const long long AA = 1LL << (2 * n), total = B * AA;
    const long long blocks = (total + 255) / 256;
    dm_init_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>(tau, total, AA);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

int dm_channel(float2 *tau, int n, long long B, float f_off, float m00, float m01, float m10, float m11, cudaStream_t s) {
    const int A = 1 << n;
    if (B > 65535) return QIDDM_EUNSUPPORTED;
    const int threads = A / 2 < 32 ? 32 : (A / 2 > 256 ? 256 : A / 2);
    dm_channel_kernel<<<dim3(A, (unsigned)B), threads, (size_t)A * sizeof(float2), s>>>(tau, n, f_off, m00, m01, m10, m11);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

int dm_transpose_conj(const float2 *src, float2 *dst, int n, long long B, cudaStream_t s) {
    const int A = 1 << n;
    if (B > 65535) return QIDDM_EUNSUPPORTED;
    const int t = (A + 31) / 32;
    dm_transpose_conj_kernel<<<dim3(t, t, (unsigned)B), 256, 0, s>>>(src, dst, A);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

int dm_readout(const float2 *tau, int n, long long B, int readout, int read_count, int read_stride, float post_scale, int clamp,
               float lo, float hi, float *out, cudaStream_t s) {
    if (readout != QIDDM_READ_PROBS && readout != QIDDM_READ_EXPVAL_Z) return QIDDM_EUNSUPPORTED;
    dm_readout_kernel<<<(unsigned)B, 256, 0, s>>>(tau, n, readout, read_count, read_stride, post_scale, clamp, lo, hi, out);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

}  // namespace qiddm
