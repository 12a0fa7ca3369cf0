// Symmetric eigensolver for the on-device replacement of the PCA-in-forward of the `*_PL*` / `differN*pca*`
// models (reference: `self.pca.fit_transform(x)` re-fit on every forward call, nn/qdense.py:456, :1429; SURVEY.md H5,
// 8f-2).  The PCA of a (m x P) batch with m <= P needs the eigen-decomposition of the m x m Gram matrix of the centred
// rows; m = batch * tau is 10..80 at the reference's batch sizes, so one CTA solves it in shared memory with the
// parallel cyclic Jacobi method (round-robin pairing: m/2 disjoint rotations per round, float64).  No host round trip
// and no status read-back, so the whole training / sampling step stays capturable in a CUDA graph.
#include <cuda_runtime.h>
#include <math.h>
#include <atomic>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

constexpr int EIGH_THREADS = 512;

// A (m x m, symmetric, row-major) -> eigenvalues in DESCENDING order and the matching eigenvectors as the columns of
// V (row-major m x m).  Shared memory: A, V (m*m doubles each), c/s per pair, diag + rank scratch.
__global__ void __launch_bounds__(EIGH_THREADS) jacobi_eigh_kernel(const double *A_batch, int m, double *evals_batch,
                                                                   double *evecs_batch, int max_sweeps) {
    // one CTA per matrix of the batch (per-group PCA: one group = the tau-ladder of one image)
    const double *A_in = A_batch + (size_t)blockIdx.x * m * m;
    double *evals = evals_batch + (size_t)blockIdx.x * m;
    double *evecs = evecs_batch + (size_t)blockIdx.x * m * m;
    extern __shared__ double sm[];
    const int ld = m | 1;                      // odd row stride (in doubles): rows start in different banks
    double *A = sm, *V = sm + (size_t)m * ld;
    double *cs = V + (size_t)m * ld;           // [2 * half]
    const int me = (m + 1) & ~1, half = me / 2;
    double *red = cs + 2 * half;               // [EIGH_THREADS / 32 + 2]
    int *pq = reinterpret_cast<int *>(red + EIGH_THREADS / 32 + 2);   // [2 * half]
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int i = tid; i < m * m; i += nt) {
        const int r = i / m, c = i - r * m;
        A[r * ld + c] = 0.5 * (A_in[i] + A_in[c * m + r]);       // symmetrise the input
        V[r * ld + c] = r == c ? 1.0 : 0.0;
    }
    __syncthreads();

    auto block_sum = [&](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < (nt >> 5); ++w) t += red[w];
        return t;
    };

    double fro = 0.0;
    for (int i = tid; i < m * m; i += nt) {
        const double v = A[(i / m) * ld + (i % m)];
        fro += v * v;
    }
    fro = block_sum(fro);

    double prev_off = 1e300;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0;
        for (int i = tid; i < m * m; i += nt) {
            const int r = i / m, c = i - r * m;
            if (r != c) off += A[r * ld + c] * A[r * ld + c];
        }
        off = block_sum(off);
        // |off|_F <= 1e-13 |A|_F, or already below 1e-10 |A|_F without further progress (rounding floor ~ m * eps)
        if (off <= 1e-26 * fro || off == 0.0 || (off <= 1e-20 * fro && off > 0.25 * prev_off)) break;
        prev_off = off;

        for (int round = 0; round < me - 1; ++round) {
            // (1) rotation of every pair of this round (round-robin: the pairs of a round are disjoint)
            for (int k = tid; k < half; k += nt) {
                int p, q;
                if (k == 0) { p = me - 1; q = round; }
                else { p = (round + k) % (me - 1); q = (round - k + (me - 1)) % (me - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < m) {
                    const double apq = A[p * ld + q];
                    if (fabs(apq) > 1e-300) {
                        const double tau = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                } else {
                    q = -1;     // padding partner of an odd m: p keeps an identity rotation
                }
                pq[2 * k] = p; pq[2 * k + 1] = q;
                cs[2 * k] = c; cs[2 * k + 1] = s;
            }
            __syncthreads();
            // (2) A <- J^T A J block by block: the 2 x 2 block (rows of pair a) x (columns of pair b) only needs the two
            // rotations, so every block is updated independently in one pass; V <- V J on its column pairs
            for (int i = tid; i < half * half; i += nt) {
                const int ka = i / half, kb = i - ka * half;
                const int pa = pq[2 * ka], qa = pq[2 * ka + 1], pb = pq[2 * kb], qb = pq[2 * kb + 1];
                const double ca = cs[2 * ka], sa = cs[2 * ka + 1], cb = cs[2 * kb], sb = cs[2 * kb + 1];
                const double b11 = A[pa * ld + pb];
                const double b12 = qb >= 0 ? A[pa * ld + qb] : 0.0;
                const double b21 = qa >= 0 ? A[qa * ld + pb] : 0.0;
                const double b22 = (qa >= 0 && qb >= 0) ? A[qa * ld + qb] : 0.0;
                const double t11 = cb * b11 - sb * b12, t12 = sb * b11 + cb * b12;
                const double t21 = cb * b21 - sb * b22, t22 = sb * b21 + cb * b22;
                A[pa * ld + pb] = ca * t11 - sa * t21;
                if (qb >= 0) A[pa * ld + qb] = ca * t12 - sa * t22;
                if (qa >= 0) A[qa * ld + pb] = sa * t11 + ca * t21;
                if (qa >= 0 && qb >= 0) A[qa * ld + qb] = sa * t12 + ca * t22;
            }
            for (int i = tid; i < half * m; i += nt) {       // consecutive threads: same row, different column pairs
                const int r = i / half, k = i - r * half;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (q < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double vp = V[r * ld + p], vq = V[r * ld + q];
                V[r * ld + p] = c * vp - s * vq;
                V[r * ld + q] = s * vp + c * vq;
            }
            __syncthreads();
        }
    }

    // descending order by rank counting (ties broken by index), then scatter
    for (int j = tid; j < m; j += nt) {
        const double lj = A[j * ld + j];
        int rank = 0;
        for (int i = 0; i < m; ++i) {
            const double li = A[i * ld + i];
            if (li > lj || (li == lj && i < j)) ++rank;
        }
        evals[rank] = lj;
        pq[j] = rank;          // m <= 2 * half
    }
    __syncthreads();
    for (int i = tid; i < m * m; i += nt) {
        const int r = i / m, c = i - r * m;
        evecs[r * m + pq[c]] = V[r * ld + c];
    }
}

}  // namespace

size_t eigh_smem_bytes(int m) {
    const int me = (m + 1) & ~1, half = me / 2;
    return (size_t)(2 * m * (m | 1) + 2 * half + EIGH_THREADS / 32 + 2) * sizeof(double) + (size_t)2 * half * sizeof(int) + 16;
}

int sym_eigh_f64(const double *A, int m, long long count, double *evals, double *evecs, cudaStream_t s) {
    if (!A || !evals || !evecs || m < 1 || count < 0 || count > 0x7fffffffLL) return QIDDM_EINVAL;
    if (count == 0) return QIDDM_OK;
    const size_t smem = eigh_smem_bytes(m);
    if (smem > 227 * 1024) return QIDDM_EUNSUPPORTED;
    // the opt-in shared-memory limit of a function is per device
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return QIDDM_ENODEVICE;
    if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63].store(true, std::memory_order_release);
    }
    // threads ~ the m/2 x m/2 block updates of a round: small matrices (m = tau = 10) use two warps, large ones all 16
    const int work = ((m + 1) / 2) * m;
    const int threads = work <= 64 ? 64 : (work <= 256 ? 128 : (work <= 1024 ? 256 : EIGH_THREADS));
    jacobi_eigh_kernel<<<(unsigned)count, threads, smem, s>>>(A, m, evals, evecs, 40);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

}  // namespace qiddm
