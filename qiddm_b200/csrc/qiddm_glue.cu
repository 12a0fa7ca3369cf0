// UNet glue around QConv2d on the device (reference nn/unet.py:28-116 runs these in float64 through
// torch.nn.Upsample(scale_factor=2, mode="bilinear") and torch.nn.BatchNorm2d): HBM-bound kernels in the tensors' own
// dtype (float64 like the reference, or float32), statistics accumulated in float64, deterministic two-stage
// reductions (no atomics).  After QConv moved to the tcgen05 path these two ops were 65 % of the UNet training step
// (library float64 kernels at 1-3 % of the HBM roofline).
#include <cuda_runtime.h>
#include <stdint.h>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// bilinear interpolation, align_corners = False (torch: src = max((dst + 0.5) * scale - 0.5, 0), i1 = min(i0 + 1, in - 1))
// ------------------------------------------------------------------------------------------------------------------
struct Lerp {
    int i0, i1;
    double l0, l1;
};
__device__ __forceinline__ Lerp lerp_of(int dst, double scale, int in_size) {
    double src = ((double)dst + 0.5) * scale - 0.5;
    if (src < 0.0) src = 0.0;
    Lerp r;
    r.i0 = (int)src;
    if (r.i0 > in_size - 1) r.i0 = in_size - 1;
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = src - (double)r.i0;
    r.l0 = 1.0 - r.l1;
    return r;
}

// (I = unsigned when every index fits 32 bits: the 64-bit div / mod per element was the instruction cost of these kernels)
template <typename T, typename I>
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const T *in, T *out, long long planes, int Hin, int Win, int Hout,
                                                           int Wout, double sh, double sw) {
    const I total = (I)(planes * Hout * Wout);
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const int ox = (int)(i % (I)Wout);
        const I t = i / (I)Wout;
        const int oy = (int)(t % (I)Hout);
        const I pl = t / (I)Hout;
        const Lerp y = lerp_of(oy, sh, Hin), x = lerp_of(ox, sw, Win);
        const T *p = in + pl * Hin * Win;
        const double v = y.l0 * (x.l0 * (double)p[y.i0 * Win + x.i0] + x.l1 * (double)p[y.i0 * Win + x.i1]) +
                         y.l1 * (x.l0 * (double)p[y.i1 * Win + x.i0] + x.l1 * (double)p[y.i1 * Win + x.i1]);
        out[i] = (T)v;
    }
}

// gather form of the transpose: every input pixel sums the output pixels that read it (deterministic, no atomics)
template <typename T, typename I>
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const T *gout, T *gin, long long planes, int Hin, int Win, int Hout,
                                                           int Wout, double sh, double sw) {
    const I total = (I)(planes * Hin * Win);
    // output rows that can touch input row iy: (iy - 1) / sh - 1 .. (iy + 1) / sh + 1
    const double ish = 1.0 / sh, isw = 1.0 / sw;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const int ix = (int)(i % (I)Win);
        const I t = i / (I)Win;
        const int iy = (int)(t % (I)Hin);
        const I pl = t / (I)Hin;
        const T *g = gout + pl * (I)(Hout * Wout);
        int oy0 = (int)(((double)iy - 1.0) * ish) - 1, oy1 = (int)(((double)iy + 1.0) * ish) + 1;
        int ox0 = (int)(((double)ix - 1.0) * isw) - 1, ox1 = (int)(((double)ix + 1.0) * isw) + 1;
        if (oy0 < 0) oy0 = 0;
        if (ox0 < 0) ox0 = 0;
        if (oy1 > Hout - 1) oy1 = Hout - 1;
        if (ox1 > Wout - 1) ox1 = Wout - 1;
        double acc = 0.0;
        // column weights once per input pixel (they do not depend on the output row): a 2x upsample has 7 candidate
        // columns and rows, 4 of each with a non-zero weight -- 16 loads instead of 49 lerp evaluations
        constexpr int WMAX = 12;
        double wxs[WMAX];
        const bool pre = ox1 - ox0 < WMAX;
        if (pre) {
#pragma unroll
            for (int k = 0; k < WMAX; ++k) {
                const int ox = ox0 + k;
                const Lerp x = lerp_of(ox <= ox1 ? ox : ox1, sw, Win);
                wxs[k] = ox <= ox1 ? (x.i0 == ix ? x.l0 : 0.0) + (x.i1 == ix ? x.l1 : 0.0) : 0.0;
            }
        }
        for (int oy = oy0; oy <= oy1; ++oy) {
            const Lerp y = lerp_of(oy, sh, Hin);
            const double wy = (y.i0 == iy ? y.l0 : 0.0) + (y.i1 == iy ? y.l1 : 0.0);
            if (wy == 0.0) continue;
            if (pre) {
                double row = 0.0;
#pragma unroll
                for (int k = 0; k < WMAX; ++k)
                    if (wxs[k] != 0.0) row += wxs[k] * (double)g[oy * Wout + ox0 + k];
                acc += wy * row;
                continue;
            }
            for (int ox = ox0; ox <= ox1; ++ox) {
                const Lerp x = lerp_of(ox, sw, Win);
                const double wx = (x.i0 == ix ? x.l0 : 0.0) + (x.i1 == ix ? x.l1 : 0.0);
                if (wx != 0.0) acc += wy * wx * (double)g[oy * Wout + ox];
            }
        }
        gin[i] = (T)acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// BatchNorm2d (NCHW), training statistics over (N, H, W) per channel
// ------------------------------------------------------------------------------------------------------------------
constexpr int BN_SPLITS = 32;      // partial sums per channel (grid.y)
constexpr int BN_THREADS = 256;

__device__ __forceinline__ void block_reduce2(double &a, double &b, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
    __syncthreads();
    if (w == 0) {
        a = l < (BN_THREADS >> 5) ? red[2 * l] : 0.0;
        b = l < (BN_THREADS >> 5) ? red[2 * l + 1] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
    }
}

// partial[c][split] = (sum x, sum x^2) over this split's share of (n, hw)
// (I = unsigned when every index fits 32 bits: the per-element div/mod is the instruction cost of these kernels)
template <typename T, typename I>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T *x, int N, int C, int HW, double *partial, int relu) {
    __shared__ double red[2 * (BN_THREADS >> 5)];
    const int c = blockIdx.x, sp = blockIdx.y;
    const I M = (I)N * (I)HW;
    double s = 0.0, q = 0.0;
    constexpr I STEP = (I)BN_SPLITS * BN_THREADS;
    I i = (I)sp * BN_THREADS + threadIdx.x;
    for (; i + 3 * STEP < M; i += 4 * STEP) {             // four loads in flight per thread
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const I iu = i + (I)u * STEP;
            const I n = iu / (I)HW;
            v[u] = (double)x[(n * (I)C + (I)c) * (I)HW + (iu - n * (I)HW)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (relu == 2 && !(v[u] > 0.0)) v[u] = 0.0;
            s += v[u];
            q += v[u] * v[u];
        }
    }
    for (; i < M; i += STEP) {
        const I n = i / (I)HW;
        const I hw = i - n * (I)HW;
        double v = (double)x[(n * (I)C + (I)c) * (I)HW + hw];
        if (relu == 2 && !(v > 0.0)) v = 0.0;           // ReLU fused in front of the normalisation (nn/unet.py:59-61)
        s += v;
        q += v * v;
    }
    block_reduce2(s, q, red);
    if (threadIdx.x == 0) {
        partial[(c * BN_SPLITS + sp) * 2] = s;
        partial[(c * BN_SPLITS + sp) * 2 + 1] = q;
    }
}

// mean / rstd per channel (+ running statistics with the unbiased variance, as torch does)
template <typename T>
__global__ void bn_finalize_kernel(const double *partial, int C, long long M, double eps, double momentum, double *save_mean,
                                   double *save_rstd, T *running_mean, T *running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < BN_SPLITS; ++i) {
        s += partial[(c * BN_SPLITS + i) * 2];
        q += partial[(c * BN_SPLITS + i) * 2 + 1];
    }
    const double mean = s / (double)M;
    double var = q / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    save_mean[c] = mean;
    save_rstd[c] = 1.0 / sqrt(var + eps);
    if (running_mean != nullptr) {
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (T)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
        running_var[c] = (T)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
    }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T *x, T *y, long long total_, int C, int HW, const double *mean,
                                                       const double *rstd, const T *gamma, const T *beta, int relu) {
    const I total = (I)total_;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const int c = (int)((i / (I)HW) % (I)C);
        const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
        double xin = (double)x[i];
        if (relu == 2 && !(xin > 0.0)) xin = 0.0;
        double v = (xin - mean[c]) * rstd[c] * g + b;
        if (relu == 1 && !(v > 0.0)) v = 0.0;           // ReLU fused behind the normalisation (nn/unet.py:95-96, :106-107)
        y[i] = (T)v;
    }
}

// partial[c][split] = (sum dy, sum dy * xhat)
template <typename T, typename I>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const T *x, const T *dy, int N, int C, int HW,
                                                                   const double *mean, const double *rstd, double *partial,
                                                                   const T *gamma, const T *beta, int relu) {
    __shared__ double red[2 * (BN_THREADS >> 5)];
    const int c = blockIdx.x, sp = blockIdx.y;
    const I M = (I)N * (I)HW;
    const double mu = mean[c], rs = rstd[c];
    const double gm = gamma ? (double)gamma[c] : 1.0, bt = beta ? (double)beta[c] : 0.0;
    double s = 0.0, q = 0.0;
    constexpr I STEP = (I)BN_SPLITS * BN_THREADS;
    I i = (I)sp * BN_THREADS + threadIdx.x;
    for (; i + STEP < M; i += 2 * STEP) {                 // two (x, dy) pairs in flight per thread
        double xv[2], gv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const I iu = i + (I)u * STEP;
            const I n = iu / (I)HW;
            const I idx = (n * (I)C + (I)c) * (I)HW + (iu - n * (I)HW);
            xv[u] = (double)x[idx];
            gv[u] = (double)dy[idx];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double xin = xv[u];
            if (relu == 2 && !(xin > 0.0)) xin = 0.0;
            const double xh = (xin - mu) * rs;
            double g = gv[u];
            if (relu == 1 && !(xh * gm + bt > 0.0)) g = 0.0;
            s += g;
            q += g * xh;
        }
    }
    for (; i < M; i += STEP) {
        const I n = i / (I)HW;
        const I hw = i - n * (I)HW;
        const I idx = (n * (I)C + (I)c) * (I)HW + hw;
        double xin = (double)x[idx];
        if (relu == 2 && !(xin > 0.0)) xin = 0.0;
        const double xh = (xin - mu) * rs;
        double g = (double)dy[idx];
        if (relu == 1 && !(xh * gm + bt > 0.0)) g = 0.0;     // threshold_backward of the fused trailing ReLU
        s += g;
        q += g * xh;
    }
    block_reduce2(s, q, red);
    if (threadIdx.x == 0) {
        partial[(c * BN_SPLITS + sp) * 2] = s;
        partial[(c * BN_SPLITS + sp) * 2 + 1] = q;
    }
}

template <typename T>
__global__ void bn_bwd_finalize_kernel(const double *partial, int C, double *sums, T *dgamma, T *dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < BN_SPLITS; ++i) {
        s += partial[(c * BN_SPLITS + i) * 2];
        q += partial[(c * BN_SPLITS + i) * 2 + 1];
    }
    sums[2 * c] = s;
    sums[2 * c + 1] = q;
    if (dbeta) dbeta[c] = (T)s;
    if (dgamma) dgamma[c] = (T)q;
}

// dx = gamma rstd (dy - sum_dy / M - xhat sum_dy_xhat / M)
template <typename T, typename I>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T *x, const T *dy, T *dx, long long total_, int C, int HW,
                                                           double inv_m, const double *mean, const double *rstd,
                                                           const T *gamma, const double *sums, const T *beta, int relu) {
    const I total = (I)total_;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const int c = (int)((i / (I)HW) % (I)C);
        const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
        const double xraw = (double)x[i];
        const double xin = (relu == 2 && !(xraw > 0.0)) ? 0.0 : xraw;
        const double xh = (xin - mean[c]) * rstd[c];
        double gy = (double)dy[i];
        if (relu == 1 && !(xh * g + b > 0.0)) gy = 0.0;
        double v = g * rstd[c] * (gy - sums[2 * c] * inv_m - xh * sums[2 * c + 1] * inv_m);
        if (relu == 2 && !(xraw > 0.0)) v = 0.0;              // threshold_backward of the fused leading ReLU
        dx[i] = (T)v;
    }
}

// MaxPool2d(kernel k, stride k) on (planes, H, W) -> (planes, H / k, W / k) (nn/unet.py:110: MaxPool2d(2, 2)); the backward
// recomputes the window's first maximum (row-major scan, strict >, as torch's kernel) instead of storing indices
template <typename T, typename I>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T *x, T *y, long long total_, int H, int W, int Ho, int Wo, int k) {
    const I total = (I)total_;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const int ox = (int)(i % (I)Wo);
        const I t = i / (I)Wo;
        const int oy = (int)(t % (I)Ho);
        const I pl = t / (I)Ho;
        const T *src = x + (pl * (I)H + (I)(oy * k)) * (I)W + (I)(ox * k);
        T m = src[0];
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
                const T v = src[(long long)dy * W + dx];
                if (v > m) m = v;
            }
        y[i] = m;
    }
}
template <typename T, typename I>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T *x, const T *gy, T *gx, long long total_, int H, int W, int Ho,
                                                          int Wo, int k) {
    const I total = (I)total_;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const int ix = (int)(i % (I)W);
        const I t = i / (I)W;
        const int iy = (int)(t % (I)H);
        const I pl = t / (I)H;
        const int oy = iy / k, ox = ix / k;
        T g = (T)0;
        if (oy < Ho && ox < Wo) {
            const T *src = x + (pl * (I)H + (I)(oy * k)) * (I)W + (I)(ox * k);
            T m = src[0];
            int arg = 0;
            for (int dy = 0; dy < k; ++dy)
                for (int dx = 0; dx < k; ++dx) {
                    const T v = src[(long long)dy * W + dx];
                    if (v > m) { m = v; arg = dy * k + dx; }
                }
            if (arg == (iy - oy * k) * k + (ix - ox * k)) g = gy[(pl * (I)Ho + (I)oy) * (I)Wo + (I)ox];
        }
        gx[i] = g;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// diffusion-step glue (src/noise.py:105-126 + src/models.py:46-67): the noise ladder written directly as the (noisy, clean)
// pair the training step consumes, and the MSE loss with its gradient in one pass
// ------------------------------------------------------------------------------------------------------------------
// level(b, t, p) = clamp(x (1 - w_t) + eps w_t, 0, 1);  noisy[(b, t)] = level(b, t + 1), clean[(b, t)] = level(b, t), t < T
template <typename T>
__global__ void __launch_bounds__(256) noise_ladder_kernel(const T *x, const float *eps, const T *w, long long n_bp, int P, int tau,
                                                           T *noisy, T *clean) {
    const int steps = tau - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bp; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / P;
        const int p = (int)(i - b * P);
        const T xv = x[i], ev = (T)eps[i];
        T prev = (T)0;
        for (int t = 0; t < tau; ++t) {
            const T wt = w[t];
            T v = xv * ((T)1 - wt) + ev * wt;
            v = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
            if (t > 0) {
                const long long o = (b * steps + (t - 1)) * P + p;
                noisy[o] = v;
                if (clean != nullptr) clean[o] = prev;
            }
            prev = v;
        }
    }
}

constexpr int MSE_BLOCKS = 148 * 8;
// d = a r + b - t1 + (t2 ? t2 : 0);  partial[block] = sum d^2;  grad = 2 a d / n
template <typename T>
__global__ void __launch_bounds__(256) mse_grad_kernel(const T *r, const T *t1, const T *t2, double a, double b, long long n,
                                                       T *grad, double *partial) {
    __shared__ double red[8];
    double acc = 0.0;
    const double k = 2.0 * a / (double)n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double d = a * (double)r[i] + b - (double)t1[i];
        if (t2 != nullptr) d += (double)t2[i];
        acc += d * d;
        grad[i] = (T)(k * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}
template <typename T>
__global__ void mse_finalize_kernel(const double *partial, int n_partial, long long n, T *loss) {   // one warp, fixed order
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) loss[0] = (T)(s / (double)n);
}

// The same loss with the target RECOMPUTED from the image and its noise draw instead of read from a materialised ladder:
// row (b, t) of pred, t < T = tau - 1:  d = a pred + b - (c0 level_t + c1 level_{t+1}),  level_t = clamp(x (1 - w_t) + eps w_t, 0, 1)
// evaluated exactly as noise_ladder_kernel does.  Goal "data": c0 = 1, c1 = 0; goal "noise": a = 0.1, b = -0.05, c0 = -1, c1 = 1.
// Reads pred + (x, eps) / T instead of pred + one or two ladder tensors: a third to a half of the pass's bytes.
template <typename T>
__global__ void __launch_bounds__(256) mse_ladder_grad_kernel(const T *r, const T *x, const float *eps, const T *w, long long batch,
                                                              int P, int steps, double a, double b, double c0, double c1, T *grad,
                                                              double *partial) {
    __shared__ double red[8];
    double acc = 0.0;
    const long long n_bp = batch * P;
    const double k = 2.0 * a / ((double)n_bp * (double)steps);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bp; i += (long long)gridDim.x * blockDim.x) {
        const long long bi = i / P;
        const int p = (int)(i - bi * P);
        const T xv = __ldg(x + i), ev = (T)__ldg(eps + i);
        const T *rp = r + bi * steps * P + p;
        T *gp = grad + bi * steps * P + p;
        const T w0 = __ldg(w);
        T prev = xv * ((T)1 - w0) + ev * w0;                             // level_0
        prev = prev < (T)0 ? (T)0 : (prev > (T)1 ? (T)1 : prev);
        for (int t0 = 0; t0 < steps; t0 += 4) {                          // rows t0 .. t0 + 3: four loads of pred in flight
            T rv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) rv[j] = t0 + j < steps ? __ldg(rp + (long long)(t0 + j) * P) : (T)0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (t0 + j < steps) {
                    const T wt = __ldg(w + t0 + j + 1);
                    T v = xv * ((T)1 - wt) + ev * wt;
                    v = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
                    const double d = a * (double)rv[j] + b - (c0 * (double)prev + c1 * (double)v);
                    acc += d * d;
                    gp[(long long)(t0 + j) * P] = (T)(k * d);
                    prev = v;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Tail of the re-upload families' training step in one pass (nn/qdense.py:642, :676 `linear_up` + src/models.py:65-67, :95-99):
//   out[r][p] = h_r . W_p + bias_p        (rows r = (image, ladder step t), K = hidden features <= 16, P pixels)
//   d = a out + b - tau_rp,  tau_rp = c0 level_t + c1 level_{t+1},  loss = mean d^2,  g = (2 a / n) d
//   dW_p = sum_r g h_r,  dbias_p = sum_r g,  dh_r = sum_p g W_p
// Neither `out` nor dL/dout (rows x P each) is materialised (the un-fused sequence makes five passes over such arrays).  The
// plain FP64 FMA rate bounds this GPU here (measured ~5 TFLOP/s), so the products are expanded: with beta_p = a bias_p + b,
//   M = sum_p W_p W_p^T, cb = sum_p beta_p W_p, sb2 = sum_p beta_p^2, H2 = sum_r h_r h_r^T, hs = sum_r h_r        (small)
//   dh_r    = (2a/n) [a M h_r + cb - V_r],             V_r  = sum_p tau_rp W_p
//   dW_p    = (2a/n) [a H2 W_p + beta_p hs - Q_p],     Q_p  = sum_r tau_rp h_r
//   dbias_p = (2a/n) [a W_p . hs + beta_p R - st_p],   st_p = sum_r tau_rp
//   n loss  = a^2 tr(M H2) + R sb2 + sum tau^2 + 2a hs . cb - 2a sum_r h_r . V_r - 2 sum_p beta_p st_p
// and only V (per image: the T + 1 ladder levels dotted with W, one warp per image) and Q / st / sum tau^2 (a thread per pixel
// walking a slice of the images) touch every (row, pixel): K + 4 and K + 8 FP64 operations instead of 2 x (2K + 12).
// ------------------------------------------------------------------------------------------------------------------
constexpr int TAIL_KMAX = 16;
constexpr int TAIL_SLICES = 192;         // image slices of tail_q_kernel (per-CTA partial sums, fixed-order reduction)
constexpr int TAIL_HSLICES = 148;        // row slices of the h moments
constexpr int TAIL_TMAX = 32;            // ladder steps held per image in tail_v_kernel

template <typename T>
__device__ __forceinline__ T ladder_level(T xv, T ev, T wt) {
    T v = xv * ((T)1 - wt) + ev * wt;
    return v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
}

// mom_w = [M (K x K) | cb (K) | sb2] (device function: run by the extra last block of tail_hmom_kernel, W staged in shared memory)
template <typename T, int K>
__device__ __forceinline__ void tail_wmom(const T *W, const T *bias, int P, double a, double b, double *mom_w, double *smem) {
    double *Ws = smem, *bs = smem + (size_t)P * K;
    for (int i = threadIdx.x; i < P * K; i += 256) Ws[i] = (double)__ldg(W + i);
    for (int i = threadIdx.x; i < P; i += 256) bs[i] = a * (bias != nullptr ? (double)__ldg(bias + i) : 0.0) + b;
    __syncthreads();
    for (int o = threadIdx.x; o < K * K + K + 1; o += 256) {
        double s0 = 0.0, s1 = 0.0;
        if (o < K * K) {
            const int i = o / K, j = o - i * K;
            int p = 0;
            for (; p + 1 < P; p += 2) {
                s0 += Ws[p * K + i] * Ws[p * K + j];
                s1 += Ws[(p + 1) * K + i] * Ws[(p + 1) * K + j];
            }
            if (p < P) s0 += Ws[p * K + i] * Ws[p * K + j];
        } else if (o < K * K + K) {
            const int k = o - K * K;
            for (int p = 0; p < P; ++p) s0 += bs[p] * Ws[p * K + k];
        } else {
            for (int p = 0; p < P; ++p) s0 += bs[p] * bs[p];
        }
        mom_w[o] = s0 + s1;
    }
}

// part_h[slice] = [H2 (K x K) | hs (K)] of a slice of the rows
template <typename T, int K>
__global__ void __launch_bounds__(256) tail_hmom_kernel(const T *h, long long rows, double *part_h, const T *W, const T *bias, int P,
                                                        double a, double b, double *mom_w) {
    extern __shared__ double tail_smem[];
    __shared__ double hsm[64 * TAIL_KMAX];
    if (blockIdx.x == gridDim.x - 1) {           // the W moments ride in the same launch
        tail_wmom<T, K>(W, bias, P, a, b, mom_w, tail_smem);
        return;
    }
    const int n_sl = gridDim.x - 1;
    const long long per = (rows + n_sl - 1) / n_sl;
    const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    constexpr int NO = K * K + K, PER = (NO + 255) / 256;          // outputs per thread (2 for K = 16)
    double s[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) s[u] = 0.0;
    for (long long c0 = r0; c0 < r1; c0 += 64) {
        const int nr = (int)(r1 - c0 < 64 ? r1 - c0 : 64);
        __syncthreads();
        for (int q = threadIdx.x; q < nr * K; q += 256) hsm[q] = (double)__ldg(h + c0 * K + q);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int o = threadIdx.x + u * 256;
            if (o < K * K) {
                const int i = o / K, j = o - i * K;
                for (int r = 0; r < nr; ++r) s[u] += hsm[r * K + i] * hsm[r * K + j];
            } else if (o < NO) {
                const int j = o - K * K;
                for (int r = 0; r < nr; ++r) s[u] += hsm[r * K + j];
            }
        }
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int o = threadIdx.x + u * 256;
        if (o < NO) part_h[(long long)blockIdx.x * NO + o] = s[u];
    }
}

// one warp per image: Vl[t] = sum_p level_t W_p for the T + 1 levels, then dh of the image's T rows and its share of sum_r h_r . V_r
template <typename T, int K>
__global__ void __launch_bounds__(256) tail_v_kernel(const T *h, const T *W, const T *x, const float *eps, const T *w,
                                                     const double *mom_w, long long batch, int P, int steps, double a, double c0,
                                                     double c1, double kk, T *dh, double *part_hv) {
    extern __shared__ double tail_smem[];
    double *Ws = tail_smem;                                  // W (P, K)
    double *Vs = tail_smem + (size_t)P * K;                  // [8 warps][TAIL_TMAX + 1][K]
    __shared__ double red[8];
    for (int i = threadIdx.x; i < P * K; i += 256) Ws[i] = (double)W[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *Vw = Vs + (size_t)warp * (TAIL_TMAX + 1) * K;
    double hv = 0.0;
    for (long long img = (long long)blockIdx.x * 8 + warp; img < batch; img += (long long)gridDim.x * 8) {
        for (int t = 0; t <= steps; ++t) {
            const T wt = __ldg(w + t);
            double acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = 0.0;
            for (int p = lane; p < P; p += 32) {
                const double lv = (double)ladder_level<T>(__ldg(x + img * P + p), (T)__ldg(eps + img * P + p), wt);
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] += lv * Ws[p * K + k];
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) Vw[t * K + k] = acc[k];
            }
        }
        __syncwarp();
        // rows of the image: lane t < steps
        for (int t = lane; t < steps; t += 32) {
            const long long r = img * steps + t;
            double hr[K];
#pragma unroll
            for (int k = 0; k < K; ++k) hr[k] = (double)__ldg(h + r * K + k);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double v = c0 * Vw[t * K + k] + c1 * Vw[(t + 1) * K + k];
                double mh = 0.0;
#pragma unroll
                for (int j = 0; j < K; ++j) mh += mom_w[k * K + j] * hr[j];
                dh[r * K + k] = (T)(kk * (a * mh + mom_w[K * K + k] - v));
                hv += hr[k] * v;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hv += __shfl_xor_sync(0xffffffffu, hv, o);
    if (lane == 0) red[warp] = hv;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        part_hv[blockIdx.x] = s;
    }
}

// a thread per pixel walking a slice of the images: Q_p = sum_r tau_rp h_r, st_p = sum_r tau_rp, and the slice's sum of tau^2
template <typename T, int K>
__global__ void __launch_bounds__(128) tail_q_kernel(const T *h, const T *x, const float *eps, const T *w, long long batch, int P,
                                                     int steps, double c0, double c1, double *part_q, double *part_t2) {
    __shared__ double hs[32 * TAIL_KMAX];        // the K hidden features of up to 32 rows of the current image
    __shared__ double red[4];
    const int p = blockIdx.x * 128 + threadIdx.x;
    const bool live = p < P;
    const long long per = (batch + gridDim.y - 1) / gridDim.y;
    const long long i0 = (long long)blockIdx.y * per, i1 = i0 + per < batch ? i0 + per : batch;
    double q[K], st = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) q[k] = 0.0;
    for (long long img = i0; img < i1; ++img) {
        const T xv = live ? __ldg(x + img * P + p) : (T)0, ev = live ? (T)__ldg(eps + img * P + p) : (T)0;
        T prev = ladder_level<T>(xv, ev, __ldg(w));
        for (int t0 = 0; t0 < steps; t0 += 32) {
            const int nt = steps - t0 < 32 ? steps - t0 : 32;
            __syncthreads();
            for (int i = threadIdx.x; i < nt * K; i += 128) hs[i] = (double)__ldg(h + (img * steps + t0) * K + i);
            __syncthreads();
            for (int t = 0; t < nt; ++t) {
                const T lv = ladder_level<T>(xv, ev, __ldg(w + t0 + t + 1));
                const double tau = c0 * (double)prev + c1 * (double)lv;
                prev = lv;
                st += tau;
                t2 += tau * tau;
#pragma unroll
                for (int k = 0; k < K; ++k) q[k] += tau * hs[t * K + k];
            }
        }
    }
    if (live) {
        double *dst = part_q + ((long long)blockIdx.y * P + p) * (K + 1);
#pragma unroll
        for (int k = 0; k < K; ++k) dst[k] = q[k];
        dst[K] = st;
    } else {
        t2 = 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t2;
    __syncthreads();
    if (threadIdx.x == 0) part_t2[blockIdx.y * gridDim.x + blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
}

// fixed-order sums of the partials and the assembly of dW (P, K) and dbias (P): a thread per (pixel, k) element (k = K: bias)
template <typename T, int K>
__global__ void __launch_bounds__(256) tail_final_kernel(const T *W, const T *bias, const double *part_h, int n_hs,
                                                         const double *part_q, int n_slices, int P, double rows, double a, double b,
                                                         double kk, T *dW, T *dbias, double *mom_h, double *stb_part) {
    __shared__ double H2[TAIL_KMAX * TAIL_KMAX + TAIL_KMAX];     // H2 then hs
    __shared__ double red[8];
    for (int o = threadIdx.x; o < K * K + K; o += 256) {
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        for (; j + 1 < n_hs; j += 2) {
            s0 += part_h[(long long)j * (K * K + K) + o];
            s1 += part_h[(long long)(j + 1) * (K * K + K) + o];
        }
        if (j < n_hs) s0 += part_h[(long long)j * (K * K + K) + o];
        H2[o] = s0 + s1;
        if (blockIdx.x == 0) mom_h[o] = s0 + s1;
    }
    __syncthreads();
    const double *hsum = H2 + K * K;
    const int e = blockIdx.x * 256 + threadIdx.x;
    double stb = 0.0;
    if (e < P * (K + 1)) {
        const int p = e / (K + 1), k = e - p * (K + 1);
        double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
        const double *src = part_q + e;
        const long long st = (long long)P * (K + 1);
        int sl = 0;
        for (; sl + 3 < n_slices; sl += 4) {
            q0 += src[sl * st];
            q1 += src[(sl + 1) * st];
            q2 += src[(sl + 2) * st];
            q3 += src[(sl + 3) * st];
        }
        for (; sl < n_slices; ++sl) q0 += src[sl * st];
        const double q = (q0 + q1) + (q2 + q3);
        const double beta = a * (bias != nullptr ? (double)bias[p] : 0.0) + b;
        double acc = 0.0;
        if (k < K) {
#pragma unroll
            for (int j = 0; j < K; ++j) acc += H2[k * K + j] * (double)W[(long long)p * K + j];
            dW[(long long)p * K + k] = (T)(kk * (a * acc + beta * hsum[k] - q));
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) acc += (double)W[(long long)p * K + j] * hsum[j];
            if (dbias != nullptr) dbias[p] = (T)(kk * (a * acc + beta * rows - q));
            stb = beta * q;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) stb += __shfl_xor_sync(0xffffffffu, stb, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = stb;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        stb_part[blockIdx.x] = s;
    }
}

template <typename T, int K>
__global__ void tail_loss_kernel(const double *mom_w, const double *mom_h, const double *part_t2, int n_t2, const double *part_hv,
                                 int n_hv, const double *stb_part, int n_stb, double rows, double n_total, double a, T *loss) {
    // one warp; every sum in a fixed order
    const int lane = threadIdx.x;
    double tr = 0.0, hcb = 0.0;
    for (int o = lane; o < K * K + K; o += 32) {
        if (o < K * K) tr += mom_w[o] * mom_h[o];        // tr(M H2): both symmetric, same index
        else hcb += mom_h[o] * mom_w[o];                 // hs . cb (same offset K * K + k in both)
    }
    double t2 = 0.0, hv = 0.0, stb = 0.0;
    for (int j = lane; j < n_t2; j += 32) t2 += part_t2[j];
    for (int j = lane; j < n_hv; j += 32) hv += part_hv[j];
    for (int j = lane; j < n_stb; j += 32) stb += stb_part[j];
    double tot = a * a * tr + 2.0 * a * hcb + t2 - 2.0 * a * hv - 2.0 * stb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) loss[0] = (T)((tot + rows * mom_w[K * K + K]) / n_total);
}

// ------------------------------------------------------------------------------------------------------------------
// Noise channels right before a computational-basis readout (nn/qdense.py:98-104, :174-180, :431-439 on default.mixed;
// src/mnist_noise.py:211-229).  A single-qubit channel applied to every wire immediately before probs() only moves
// population: PhaseShift / PhaseDamping are diagonal (no effect), AmplitudeDamping(g) maps (p0, p1) -> (p0 + g p1,
// (1 - g) p1), DepolarizingChannel(q) (Kraus sqrt(1-q) I, sqrt(q/3) X, Y, Z) flips the bit with probability 2q/3.  So
// the density-matrix simulation reduces EXACTLY to one 2 x 2 column-stochastic matrix M per wire acting on the
// probability vector: p' = (M x ... x M) p, n butterfly stages in shared memory, one CTA per instance.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) prob_channel_kernel(const T *p_in, T *p_out, int n, double m00, double m01, double m10,
                                                           double m11) {
    extern __shared__ double ch_sm[];
    const int A = 1 << n;
    const T *src = p_in + (size_t)blockIdx.x * A;
    for (int k = threadIdx.x; k < A; k += blockDim.x) ch_sm[k] = (double)src[k];
    __syncthreads();
    for (int b = 0; b < n; ++b) {
        for (int i = threadIdx.x; i < A / 2; i += blockDim.x) {
            const int lo = i & ((1 << b) - 1);
            const int k0 = ((i >> b) << (b + 1)) | lo, k1 = k0 | (1 << b);
            const double a = ch_sm[k0], c = ch_sm[k1];
            ch_sm[k0] = m00 * a + m01 * c;
            ch_sm[k1] = m10 * a + m11 * c;
        }
        __syncthreads();
    }
    T *dst = p_out + (size_t)blockIdx.x * A;
    for (int k = threadIdx.x; k < A; k += blockDim.x) dst[k] = (T)ch_sm[k];
}

// ------------------------------------------------------------------------------------------------------------------
// The forward the reference's `_QConv2d_FAST` ACTUALLY executes (nn/qconv.py:71-90; SURVEY.md H1: the QNode is never
// called): unfold -> +0.1 -> * F * 0.5 -> clamp(0, 1) -> [:, ::2] -> [:, :out_channels], F = C kh kw.  Output channel j is
// the patch feature 2j = (ch, ky, kx) -- one pixel per output element, so the map and its gradient are gathers.
// ------------------------------------------------------------------------------------------------------------------
struct RefMapGeom {
    int C, H, W, kh, kw, ph, pw, Hout, Wout, n_ch_out;
};
template <typename T>
__global__ void __launch_bounds__(256) qconv_refmap_fwd_kernel(const T *img, T *out, long long total, RefMapGeom g) {
    const int kk = g.kh * g.kw;
    const T F = (T)(g.C * kk);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.Wout);
        long long t = i / g.Wout;
        const int y = (int)(t % g.Hout);
        t /= g.Hout;
        const int j = (int)(t % g.n_ch_out);
        const long long b = t / g.n_ch_out;
        const int f = 2 * j, ch = f / kk, r = f - ch * kk, ky = r / g.kw, kx = r - ky * g.kw;
        const int iy = y + ky - g.ph, ix = x + kx - g.pw;
        T v = (T)0;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = img[((b * g.C + ch) * g.H + iy) * g.W + ix];
        v = (v + (T)0.1) * F * (T)0.5;
        out[i] = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
    }
}
// grad_img[b,ch,iy,ix] = mask(pixel) * F/2 * sum over the retained features (ch,ky,kx) = 2j of grad_out[b,j,iy-ky+ph,ix-kx+pw]
template <typename T>
__global__ void __launch_bounds__(256) qconv_refmap_bwd_kernel(const T *img, const T *gout, T *gimg, long long total,
                                                               RefMapGeom g) {
    const int kk = g.kh * g.kw;
    const T F = (T)(g.C * kk);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % g.W);
        long long t = i / g.W;
        const int iy = (int)(t % g.H);
        t /= g.H;
        const int ch = (int)(t % g.C);
        const long long b = t / g.C;
        T acc = (T)0;
        for (int r = 0; r < kk; ++r) {
            const int f = ch * kk + r;
            if ((f & 1) || (f >> 1) >= g.n_ch_out) continue;
            const int ky = r / g.kw, kx = r - ky * g.kw;
            const int y = iy - ky + g.ph, x = ix - kx + g.pw;
            if (y < 0 || y >= g.Hout || x < 0 || x >= g.Wout) continue;
            acc += gout[((b * g.n_ch_out + (f >> 1)) * g.Hout + y) * g.Wout + x];
        }
        const T v = (img[i] + (T)0.1) * F * (T)0.5;
        // zero-padding pixels feed outputs too, but carry no gradient to the image
        gimg[i] = (v >= (T)0 && v <= (T)1) ? acc * F * (T)0.5 : (T)0;
    }
}

inline unsigned ew_grid(long long total) {
    const long long b = (total + 255) / 256;
    return (unsigned)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

template <typename T>
int upsample_impl(const void *in, void *out, bool backward, long long planes, int Hin, int Win, int Hout, int Wout, double sh,
                  double sw, cudaStream_t s) {
    const bool small = planes * Hout * Wout < (1LL << 31) && planes * Hin * Win < (1LL << 31);
    const T *src = reinterpret_cast<const T *>(in);
    T *dst = reinterpret_cast<T *>(out);
    if (backward) {  // `in` = grad_out (planes, Hout, Wout), `out` = grad_in (planes, Hin, Win)
        // (measured: the gather kernel is SLOWER with 32-bit indices, 85 vs 57 us per launch at 640 x 16 x 28 x 28 -- kept on 64-bit)
        upsample_bwd_kernel<T, long long><<<ew_grid(planes * Hin * Win), 256, 0, s>>>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw);
    } else {
        if (small) upsample_fwd_kernel<T, unsigned><<<ew_grid(planes * Hout * Wout), 256, 0, s>>>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw);
        else upsample_fwd_kernel<T, long long><<<ew_grid(planes * Hout * Wout), 256, 0, s>>>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

template <typename T, typename I>
int bn_fwd_impl(const void *x, void *y, int N, int C, int HW, const void *gamma, const void *beta, double *save_mean,
                double *save_rstd, void *running_mean, void *running_var, double momentum, double eps, double *ws,
                cudaStream_t s, int relu) {
    const long long M = (long long)N * HW, total = M * C;
    bn_stats_kernel<T, I><<<dim3(C, BN_SPLITS), BN_THREADS, 0, s>>>(reinterpret_cast<const T *>(x), N, C, HW, ws, relu);
    bn_finalize_kernel<T><<<(C + 127) / 128, 128, 0, s>>>(ws, C, M, eps, momentum, save_mean, save_rstd,
                                                         reinterpret_cast<T *>(running_mean), reinterpret_cast<T *>(running_var));
    bn_apply_kernel<T, I><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const T *>(x), reinterpret_cast<T *>(y), total, C, HW,
                                                     save_mean, save_rstd, reinterpret_cast<const T *>(gamma),
                                                     reinterpret_cast<const T *>(beta), relu);
    count_launch(3);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

template <typename T, typename I>
int bn_bwd_impl(const void *x, const void *dy, void *dx, int N, int C, int HW, const void *gamma, const double *save_mean,
                const double *save_rstd, void *dgamma, void *dbeta, double *ws, cudaStream_t s, const void *beta, int relu) {
    const long long M = (long long)N * HW, total = M * C;
    double *sums = ws + (size_t)C * BN_SPLITS * 2;
    bn_bwd_reduce_kernel<T, I><<<dim3(C, BN_SPLITS), BN_THREADS, 0, s>>>(reinterpret_cast<const T *>(x), reinterpret_cast<const T *>(dy),
                                                                      N, C, HW, save_mean, save_rstd, ws,
                                                                      reinterpret_cast<const T *>(gamma), reinterpret_cast<const T *>(beta), relu);
    bn_bwd_finalize_kernel<T><<<(C + 127) / 128, 128, 0, s>>>(ws, C, sums, reinterpret_cast<T *>(dgamma), reinterpret_cast<T *>(dbeta));
    int launches = 2;
    if (dx != nullptr) {
        bn_bwd_apply_kernel<T, I><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const T *>(x), reinterpret_cast<const T *>(dy),
                                                             reinterpret_cast<T *>(dx), total, C, HW, 1.0 / (double)M, save_mean,
                                                             save_rstd, reinterpret_cast<const T *>(gamma), sums,
                                                             reinterpret_cast<const T *>(beta), relu);
        ++launches;
    }
    count_launch(launches);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

template <typename T>
int ladder_impl(const void *x, const float *eps, const void *w, long long batch, int P, int tau, void *noisy, void *clean,
                cudaStream_t s) {
    noise_ladder_kernel<T><<<ew_grid(batch * P), 256, 0, s>>>(reinterpret_cast<const T *>(x), eps, reinterpret_cast<const T *>(w),
                                                              batch * P, P, tau, reinterpret_cast<T *>(noisy), reinterpret_cast<T *>(clean));
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}
template <typename T>
int mse_impl(const void *r, const void *t1, const void *t2, double a, double b, long long n, void *grad, void *loss, double *ws,
             cudaStream_t s) {
    const int blocks = (int)((n + 255) / 256 < MSE_BLOCKS ? (n + 255) / 256 : MSE_BLOCKS);
    mse_grad_kernel<T><<<blocks, 256, 0, s>>>(reinterpret_cast<const T *>(r), reinterpret_cast<const T *>(t1),
                                              reinterpret_cast<const T *>(t2), a, b, n, reinterpret_cast<T *>(grad), ws);
    mse_finalize_kernel<T><<<1, 32, 0, s>>>(ws, blocks, n, reinterpret_cast<T *>(loss));
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

template <typename T>
int mse_ladder_impl(const void *r, const void *x, const float *eps, const void *w, long long batch, int P, int tau, double a,
                    double b, double c0, double c1, void *grad, void *loss, double *ws, cudaStream_t s) {
    const long long n_bp = batch * P;
    const int blocks = (int)((n_bp + 255) / 256 < MSE_BLOCKS ? (n_bp + 255) / 256 : MSE_BLOCKS);
    mse_ladder_grad_kernel<T><<<blocks, 256, 0, s>>>(reinterpret_cast<const T *>(r), reinterpret_cast<const T *>(x), eps,
                                                     reinterpret_cast<const T *>(w), batch, P, tau - 1, a, b, c0, c1,
                                                     reinterpret_cast<T *>(grad), ws);
    mse_finalize_kernel<T><<<1, 32, 0, s>>>(ws, blocks, n_bp * (tau - 1), reinterpret_cast<T *>(loss));
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

}  // namespace

int prob_channel(const void *p_in, void *p_out, int dtype, long long batch, int n, double m00, double m01, double m10, double m11,
                 cudaStream_t s) {
    if (!p_in || !p_out || batch < 0 || n < 1 || n > QIDDM_MAX_QUBITS || batch > 0x7fffffffLL) return QIDDM_EINVAL;
    if (batch == 0) return QIDDM_OK;
    const size_t smem = ((size_t)1 << n) * sizeof(double);
    const int threads = (1 << n) / 2 < 256 ? ((1 << n) / 2 < 32 ? 32 : (1 << n) / 2) : 256;
    if (dtype == QIDDM_DTYPE_F64)
        prob_channel_kernel<double><<<(unsigned)batch, threads, smem, s>>>(reinterpret_cast<const double *>(p_in),
                                                                            reinterpret_cast<double *>(p_out), n, m00, m01, m10, m11);
    else if (dtype == QIDDM_DTYPE_F32)
        prob_channel_kernel<float><<<(unsigned)batch, threads, smem, s>>>(reinterpret_cast<const float *>(p_in),
                                                                           reinterpret_cast<float *>(p_out), n, m00, m01, m10, m11);
    else
        return QIDDM_EINVAL;
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

// FP32 FMA-pipe probe (the roofline denominator of the gate kernels, which MEASURED_PEAKS.json does not hold): every thread
// runs 8 independent packed-FMA chains, `iters` rounds of 16 fma.rn.f32x2 = 64 flop per round; no memory traffic.
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *sink, int iters, float a, float b) {
    float2 x[8], y[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, (float)i); y[i] = make_float2(i * 0.5f, 1.f); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __ffma2_rn(a2, x[i], y[i]); y[i] = __ffma2_rn(b2, y[i], x[i]); }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + y[i].x + y[i].y;
    if (s == 1.2345e-30f) sink[0] = s;          // never true: keeps the chains alive without a store per thread
}
int probe_fp32_fma(int iters, float *sink, double *flops, cudaStream_t s) {
    if (iters < 1 || !sink) return QIDDM_EINVAL;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return QIDDM_ENODEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;
    fp32_probe_kernel<<<grid, 256, 0, s>>>(sink, iters, 0.999f, 1.001f);
    count_launch();
    if (flops) *flops = (double)grid * 256.0 * (double)iters * 64.0;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

int qconv_reference_map(const void *img, const void *grad_out, void *out, int dtype, bool backward, long long n_images, int C,
                        int H, int W, int kh, int kw, int ph, int pw, int n_ch_out, cudaStream_t s) {
    if (!img || !out || n_images < 0 || C < 1 || H < 1 || W < 1 || kh < 1 || kw < 1 || ph < 0 || pw < 0) return QIDDM_EINVAL;
    if (backward && !grad_out) return QIDDM_EINVAL;
    RefMapGeom g{C, H, W, kh, kw, ph, pw, H + 2 * ph - kh + 1, W + 2 * pw - kw + 1, n_ch_out};
    if (g.Hout < 1 || g.Wout < 1 || n_ch_out < 1 || 2 * (n_ch_out - 1) >= C * kh * kw) return QIDDM_EINVAL;
    if (n_images == 0) return QIDDM_OK;
    const long long total = backward ? n_images * C * H * W : n_images * n_ch_out * g.Hout * g.Wout;
    if (dtype == QIDDM_DTYPE_F64) {
        if (backward)
            qconv_refmap_bwd_kernel<double><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const double *>(img),
                                                                          reinterpret_cast<const double *>(grad_out),
                                                                          reinterpret_cast<double *>(out), total, g);
        else
            qconv_refmap_fwd_kernel<double><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const double *>(img),
                                                                          reinterpret_cast<double *>(out), total, g);
    } else if (dtype == QIDDM_DTYPE_F32) {
        if (backward)
            qconv_refmap_bwd_kernel<float><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const float *>(img),
                                                                         reinterpret_cast<const float *>(grad_out),
                                                                         reinterpret_cast<float *>(out), total, g);
        else
            qconv_refmap_fwd_kernel<float><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const float *>(img),
                                                                         reinterpret_cast<float *>(out), total, g);
    } else {
        return QIDDM_EINVAL;
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

size_t mse_ws_bytes() { return (size_t)MSE_BLOCKS * sizeof(double); }

int noise_ladder(const void *x, const float *eps, const void *w, int dtype, long long batch, int P, int tau, void *noisy,
                 void *clean, cudaStream_t s) {
    if (!x || !eps || !w || !noisy || batch < 0 || P < 1 || tau < 2) return QIDDM_EINVAL;       // clean may be NULL
    if (batch == 0) return QIDDM_OK;
    if (dtype == QIDDM_DTYPE_F64) return ladder_impl<double>(x, eps, w, batch, P, tau, noisy, clean, s);
    if (dtype == QIDDM_DTYPE_F32) return ladder_impl<float>(x, eps, w, batch, P, tau, noisy, clean, s);
    return QIDDM_EINVAL;
}

int mse_loss_grad(const void *r, const void *t1, const void *t2, int dtype, double a, double b, long long n, void *grad,
                  void *loss, void *ws, cudaStream_t s) {
    if (!r || !t1 || !grad || !loss || !ws || n < 1) return QIDDM_EINVAL;
    if (dtype == QIDDM_DTYPE_F64) return mse_impl<double>(r, t1, t2, a, b, n, grad, loss, reinterpret_cast<double *>(ws), s);
    if (dtype == QIDDM_DTYPE_F32) return mse_impl<float>(r, t1, t2, a, b, n, grad, loss, reinterpret_cast<double *>(ws), s);
    return QIDDM_EINVAL;
}

int mse_ladder_loss_grad(const void *r, const void *x, const float *eps, const void *w, int dtype, long long batch, int P, int tau,
                         double a, double b, double c0, double c1, void *grad, void *loss, void *ws, cudaStream_t s) {
    if (!r || !x || !eps || !w || !grad || !loss || !ws || batch < 1 || P < 1 || tau < 2) return QIDDM_EINVAL;
    if (dtype == QIDDM_DTYPE_F64)
        return mse_ladder_impl<double>(r, x, eps, w, batch, P, tau, a, b, c0, c1, grad, loss, reinterpret_cast<double *>(ws), s);
    if (dtype == QIDDM_DTYPE_F32)
        return mse_ladder_impl<float>(r, x, eps, w, batch, P, tau, a, b, c0, c1, grad, loss, reinterpret_cast<double *>(ws), s);
    return QIDDM_EINVAL;
}

size_t linear_up_mse_ws_bytes(int P, int K) {
    const size_t pb = (size_t)(P + 127) / 128;
    return ((size_t)TAIL_SLICES * P * (K + 1) + TAIL_SLICES * pb + (size_t)K * K + K + 1 + (size_t)(TAIL_HSLICES + 1) * (K * K + K) +
            148 * 4 + ((size_t)P * (K + 1) + 255) / 256 + 64) * sizeof(double) + 256;
}

namespace {
template <typename T, int K>
int tail_impl(const void *h, const void *W, const void *bias, const void *x, const float *eps, const void *w, long long batch, int P,
              int tau, double a, double b, double c0, double c1, void *loss, void *dW, void *dbias, void *dh, void *ws, cudaStream_t s) {
    const int steps = tau - 1, pb = (P + 127) / 128;
    if (steps > TAIL_TMAX) return QIDDM_EUNSUPPORTED;
    const long long rows = batch * steps;
    const int slices = (int)(batch < TAIL_SLICES ? batch : TAIL_SLICES);
    const int hslices = (int)(rows < TAIL_HSLICES ? rows : TAIL_HSLICES);
    const int vgrid = (int)((batch + 7) / 8 < 148 * 4 ? (batch + 7) / 8 : 148 * 4);
    const int fgrid = (P * (K + 1) + 255) / 256;
    double *part_q = reinterpret_cast<double *>(ws);
    double *part_t2 = part_q + (size_t)TAIL_SLICES * P * (K + 1);
    double *mom_w = part_t2 + (size_t)TAIL_SLICES * pb;
    double *part_h = mom_w + (K * K + K + 1);
    double *mom_h = part_h + (size_t)TAIL_HSLICES * (K * K + K);
    double *part_hv = mom_h + (K * K + K);
    double *stb_part = part_hv + 148 * 4;
    const T *hp = reinterpret_cast<const T *>(h), *Wp = reinterpret_cast<const T *>(W), *bp = reinterpret_cast<const T *>(bias);
    const T *xp = reinterpret_cast<const T *>(x), *wp = reinterpret_cast<const T *>(w);
    const double kk = 2.0 * a / ((double)rows * (double)P);
    const size_t smem_w = (size_t)P * (K + 1) * sizeof(double);
    const size_t smem_v = ((size_t)P * K + (size_t)8 * (TAIL_TMAX + 1) * K) * sizeof(double);
    auto km = tail_hmom_kernel<T, K>;
    auto kv = tail_v_kernel<T, K>;
    if (smem_w > 40 * 1024 && cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w) != cudaSuccess)
        return QIDDM_EUNSUPPORTED;
    if (smem_v > 48 * 1024 && cudaFuncSetAttribute(kv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v) != cudaSuccess)
        return QIDDM_EUNSUPPORTED;
    T *dh_out = reinterpret_cast<T *>(dh);
    if (dh_out == nullptr) return QIDDM_EUNSUPPORTED;        // V also feeds the loss: grad_h is always written
    km<<<hslices + 1, 256, smem_w, s>>>(hp, rows, part_h, Wp, bp, P, a, b, mom_w);
    tail_q_kernel<T, K><<<dim3(pb, slices), 128, 0, s>>>(hp, xp, eps, wp, batch, P, steps, c0, c1, part_q, part_t2);
    kv<<<vgrid, 256, smem_v, s>>>(hp, Wp, xp, eps, wp, mom_w, batch, P, steps, a, c0, c1, kk, dh_out, part_hv);
    tail_final_kernel<T, K><<<fgrid, 256, 0, s>>>(Wp, bp, part_h, hslices, part_q, slices, P, (double)rows, a, b, kk,
                                                   reinterpret_cast<T *>(dW), reinterpret_cast<T *>(dbias), mom_h, stb_part);
    tail_loss_kernel<T, K><<<1, 32, 0, s>>>(mom_w, mom_h, part_t2, slices * pb, part_hv, vgrid, stb_part, fgrid, (double)rows,
                                            (double)rows * P, a, reinterpret_cast<T *>(loss));
    count_launch(5);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}
template <typename T>
int tail_k(int K, const void *h, const void *W, const void *bias, const void *x, const float *eps, const void *w, long long batch,
           int P, int tau, double a, double b, double c0, double c1, void *loss, void *dW, void *dbias, void *dh, void *ws,
           cudaStream_t s) {
#define QIDDM_TAIL(K_) case K_: return tail_impl<T, K_>(h, W, bias, x, eps, w, batch, P, tau, a, b, c0, c1, loss, dW, dbias, dh, ws, s)
    switch (K) {
        QIDDM_TAIL(1); QIDDM_TAIL(2); QIDDM_TAIL(3); QIDDM_TAIL(4); QIDDM_TAIL(5); QIDDM_TAIL(6); QIDDM_TAIL(7); QIDDM_TAIL(8);
        QIDDM_TAIL(9); QIDDM_TAIL(10); QIDDM_TAIL(11); QIDDM_TAIL(12); QIDDM_TAIL(13); QIDDM_TAIL(14); QIDDM_TAIL(15); QIDDM_TAIL(16);
    }
#undef QIDDM_TAIL
    return QIDDM_EUNSUPPORTED;
}
}  // namespace

int linear_up_mse_step(const void *h, const void *W, const void *bias, const void *x, const float *eps, const void *w, int dtype,
                       long long batch, int P, int tau, int K, double a, double b, double c0, double c1, void *loss, void *dW,
                       void *dbias, void *dh, void *ws, cudaStream_t s) {
    if (!h || !W || !x || !eps || !w || !loss || !dW || !ws || batch < 1 || P < 1 || tau < 2) return QIDDM_EINVAL;
    if (K < 1 || K > TAIL_KMAX || ((size_t)P * K + (size_t)8 * (TAIL_TMAX + 1) * K) * sizeof(double) > 200 * 1024 || !dh)
        return QIDDM_EUNSUPPORTED;       // W and the per-warp level sums must fit in shared memory; grad_h is always written
    if (dtype == QIDDM_DTYPE_F64) return tail_k<double>(K, h, W, bias, x, eps, w, batch, P, tau, a, b, c0, c1, loss, dW, dbias, dh, ws, s);
    if (dtype == QIDDM_DTYPE_F32) return tail_k<float>(K, h, W, bias, x, eps, w, batch, P, tau, a, b, c0, c1, loss, dW, dbias, dh, ws, s);
    return QIDDM_EINVAL;
}

size_t batchnorm_ws_bytes(int C) { return ((size_t)C * BN_SPLITS * 2 + (size_t)C * 2) * sizeof(double); }

int upsample_bilinear(const void *in, void *out, int dtype, bool backward, long long planes, int Hin, int Win, int Hout, int Wout,
                      double scale_h, double scale_w, cudaStream_t s) {
    if (!in || !out || planes < 0 || Hin < 1 || Win < 1 || Hout < 1 || Wout < 1) return QIDDM_EINVAL;
    if (planes == 0) return QIDDM_OK;
    if (dtype == QIDDM_DTYPE_F64) return upsample_impl<double>(in, out, backward, planes, Hin, Win, Hout, Wout, scale_h, scale_w, s);
    if (dtype == QIDDM_DTYPE_F32) return upsample_impl<float>(in, out, backward, planes, Hin, Win, Hout, Wout, scale_h, scale_w, s);
    return QIDDM_EINVAL;
}

int batchnorm_forward(const void *x, void *y, int dtype, int N, int C, int HW, const void *gamma, const void *beta,
                      double *save_mean, double *save_rstd, void *running_mean, void *running_var, double momentum, double eps,
                      void *ws, cudaStream_t s, int relu) {
    if (!x || !y || !save_mean || !save_rstd || !ws || N < 1 || C < 1 || HW < 1 || relu < 0 || relu > 2) return QIDDM_EINVAL;
    if ((running_mean == nullptr) != (running_var == nullptr)) return QIDDM_EINVAL;
    const bool small = (long long)N * C * HW < (1LL << 31);
    double *w = reinterpret_cast<double *>(ws);
#define QIDDM_BN_FWD(T, I) bn_fwd_impl<T, I>(x, y, N, C, HW, gamma, beta, save_mean, save_rstd, running_mean, running_var, momentum, eps, w, s, relu)
    if (dtype == QIDDM_DTYPE_F64) return small ? QIDDM_BN_FWD(double, unsigned) : QIDDM_BN_FWD(double, long long);
    if (dtype == QIDDM_DTYPE_F32) return small ? QIDDM_BN_FWD(float, unsigned) : QIDDM_BN_FWD(float, long long);
#undef QIDDM_BN_FWD
    return QIDDM_EINVAL;
}

int batchnorm_backward(const void *x, const void *dy, void *dx, int dtype, int N, int C, int HW, const void *gamma,
                       const double *save_mean, const double *save_rstd, void *dgamma, void *dbeta, void *ws, cudaStream_t s,
                       const void *beta, int relu) {
    if (!x || !dy || !save_mean || !save_rstd || !ws || N < 1 || C < 1 || HW < 1 || relu < 0 || relu > 2) return QIDDM_EINVAL;
    const bool small = (long long)N * C * HW < (1LL << 31);
    double *w = reinterpret_cast<double *>(ws);
#define QIDDM_BN_BWD(T, I) bn_bwd_impl<T, I>(x, dy, dx, N, C, HW, gamma, save_mean, save_rstd, dgamma, dbeta, w, s, beta, relu)
    if (dtype == QIDDM_DTYPE_F64) return small ? QIDDM_BN_BWD(double, unsigned) : QIDDM_BN_BWD(double, long long);
    if (dtype == QIDDM_DTYPE_F32) return small ? QIDDM_BN_BWD(float, unsigned) : QIDDM_BN_BWD(float, long long);
#undef QIDDM_BN_BWD
    return QIDDM_EINVAL;
}

int maxpool2d(const void *x, const void *gy, void *out, int dtype, bool backward, long long planes, int H, int W, int k,
              cudaStream_t s) {
    if (!x || !out || planes < 0 || H < 1 || W < 1 || k < 1 || (backward && !gy)) return QIDDM_EINVAL;
    const int Ho = H / k, Wo = W / k;
    if (Ho < 1 || Wo < 1) return QIDDM_EINVAL;
    if (planes == 0) return QIDDM_OK;
    const long long total = backward ? planes * H * W : planes * Ho * Wo;
    const bool small = planes * H * W < (1LL << 31);
#define QIDDM_POOL(T, I)                                                                                                            \
    do {                                                                                                                            \
        if (backward) maxpool_bwd_kernel<T, I><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const T *>(x), reinterpret_cast<const T *>(gy), \
                                                                              reinterpret_cast<T *>(out), total, H, W, Ho, Wo, k); \
        else maxpool_fwd_kernel<T, I><<<ew_grid(total), 256, 0, s>>>(reinterpret_cast<const T *>(x), reinterpret_cast<T *>(out), total, H, W, \
                                                                     Ho, Wo, k);                                                   \
    } while (0)
    if (dtype == QIDDM_DTYPE_F64) {
        if (small) QIDDM_POOL(double, unsigned); else QIDDM_POOL(double, long long);
    } else if (dtype == QIDDM_DTYPE_F32) {
        if (small) QIDDM_POOL(float, unsigned); else QIDDM_POOL(float, long long);
    } else {
        return QIDDM_EINVAL;
    }
#undef QIDDM_POOL
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

}  // namespace qiddm
